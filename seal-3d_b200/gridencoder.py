"""Host-side mirror of the reference's ``gridencoder/grid.py`` (``grid_encode`` / ``GridEncoder``).

Same constructor arguments, parameter names (``embeddings``, ``offsets``: reference checkpoints load
unchanged) and autocast behaviour (fp16 table when autocast is on and C is even).  What changes:
the fp16 copy of the table is a persistent shadow refreshed only when the parameter was modified
(the reference re-casts all 12.2 M entries on every call, grid.py:43-44).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


class _grid_encode(Function):
    """grid.py:24-90.  inputs [B,D] float in [0,1]; embeddings [sO,C]; offsets int32 [L+1] -> [B, L*C]."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
        inputs = inputs.contiguous().float()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        if embeddings.dtype not in (torch.float32, torch.float16):
            raise _lib.S3DError("embeddings must be float32 or float16")
        embeddings = embeddings.contiguous()
        dt = 0 if embeddings.dtype == torch.float32 else 1
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
        _lib.check_cuda(inputs, embeddings, offsets)
        _lib.call("s3d_grid_encode_forward", inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, int(gridtype),
                  int(bool(align_corners)), int(interpolation), dt)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, int(gridtype), int(interpolation), dt)
        ctx.align_corners = bool(align_corners)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, dt = ctx.dims
        grad = grad.to(embeddings.dtype).view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
        _lib.call("s3d_grid_encode_backward", grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                  grad_inputs, gridtype, int(ctx.align_corners), interpolation, dt)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


def grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
    return _grid_encode.apply(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs, gridtype,
                              align_corners, interpolation)


class _to_half_shadow(Function):
    """fp32 parameter -> cached fp16 copy; gradient flows back to the fp32 parameter."""

    @staticmethod
    def forward(ctx, param, shadow):
        return shadow

    @staticmethod
    def backward(ctx, g):
        return g.float(), None


class GridEncoder(nn.Module):
    """grid.py:96-185."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False, interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.log2_hashmap_size, self.base_resolution = per_level_scale, log2_hashmap_size, base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id = gridtype, _gridtype_to_id[gridtype]
        self.interpolation, self.interp_id = interpolation, _interp_to_id[interpolation]
        self.align_corners = align_corners
        offsets, offset = [], 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            n = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            n = int(np.ceil(n / 8) * 8)
            offsets.append(offset)
            offset += n
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()
        self._shadow, self._shadow_version = None, -1

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def half_table(self):
        """persistent fp16 shadow of the table, refreshed only after the parameter changed"""
        ext = getattr(self, "external_shadow", None)
        if ext is not None:  # kept fresh by the fused Adam kernel (trainer.ParamArena)
            return ext
        ver = self.embeddings._version
        if self._shadow is None or self._shadow_version != ver or self._shadow.device != self.embeddings.device:
            if self._shadow is None or self._shadow.device != self.embeddings.device:
                self._shadow = torch.empty_like(self.embeddings, dtype=torch.float16)
            _lib.call("s3d_cast_f32_to_f16", self.embeddings.detach().contiguous(), self._shadow, self.embeddings.numel())
            self._shadow_version = ver
        return self._shadow

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        emb = self.embeddings
        if torch.is_autocast_enabled() and self.level_dim % 2 == 0:
            emb = _to_half_shadow.apply(self.embeddings, self.half_table())
        out = grid_encode(inputs, emb, self.offsets, self.per_level_scale, self.base_resolution, inputs.requires_grad,
                          self.gridtype_id, self.align_corners, self.interp_id)
        return out.view(prefix + [self.output_dim])

    @torch.no_grad()
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """grid.py:163-185: adds the TV-regulariser gradient into embeddings.grad."""
        D, C, L = self.input_dim, self.embeddings.shape[1], self.offsets.shape[0] - 1
        S, H = float(np.log2(self.per_level_scale)), self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, D, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, D)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        _lib.call("s3d_grad_total_variation", inputs.contiguous().float(), self.embeddings.detach(), self.embeddings.grad, self.offsets,
                  float(weight), B, D, C, L, S, H, self.gridtype_id, int(self.align_corners), 0)
