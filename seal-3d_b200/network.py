"""NGP field used by Seal-3D (mirror of ``nerf/network.py``: hash grid -> sigma MLP 32-64-16 with
trunc_exp; SH(4) + geo(15) + second hash grid(32) -> colour MLP 63-64-64-3 with sigmoid), on top of
the renderer mirror.  Parameter names and shapes equal the reference's state dict (SURVEY.md
appendix B) so reference checkpoints load with ``load_state_dict``.

Two execution paths share the same parameters:
  * ``forward`` / ``density`` / ``color`` -- op-by-op like the reference (our grid / SH kernels plus
    torch ``F.linear`` GEMMs), differentiable through autograd;
  * ``fused_*`` (seal3d_b200.fused) -- the single-kernel tcgen05 field used by the distillation trainer.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from .gridencoder import GridEncoder
from .shencoder import SHEncoder
from .renderer import NeRFRenderer


class _trunc_exp(Function):
    """activation.py:5-17: exp forward (fp32), backward multiplies by exp(clamp(x, -15, 15))."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _trunc_exp.apply


def get_encoder(encoding, input_dim=3, degree=4, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=2048, align_corners=False, **kwargs):
    """encoding.py:45-77 for the encodings on the hot path."""
    if encoding == "sphere_harmonics":
        enc = SHEncoder(input_dim=input_dim, degree=degree)
    elif encoding in ("hashgrid", "tiledgrid"):
        enc = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                          log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                          gridtype="hash" if encoding == "hashgrid" else "tiled", align_corners=align_corners)
    elif encoding == "frequency":
        from .freqencoder import FreqEncoder
        enc = FreqEncoder(input_dim=input_dim, degree=kwargs.get("multires", 6))
    else:
        raise NotImplementedError("Unknown encoding mode, choose from [frequency, sphere_harmonics, hashgrid, tiledgrid]")
    return enc, enc.output_dim


class NeRFNetwork(NeRFRenderer):
    def __init__(self, encoding="hashgrid", encoding_dir="sphere_harmonics", num_layers=2, hidden_dim=64, geo_feat_dim=15,
                 num_layers_color=3, hidden_dim_color=64, bound=1, log2_hashmap_size=19, **kwargs):
        super().__init__(bound, **kwargs)
        self.num_layers, self.hidden_dim, self.geo_feat_dim = num_layers, hidden_dim, geo_feat_dim
        self.encoder, self.in_dim = get_encoder(encoding, desired_resolution=2048 * bound, log2_hashmap_size=log2_hashmap_size)
        self.sigma_net = nn.ModuleList([
            nn.Linear(self.in_dim if l == 0 else hidden_dim, 1 + geo_feat_dim if l == num_layers - 1 else hidden_dim, bias=False)
            for l in range(num_layers)])
        self.num_layers_color, self.hidden_dim_color = num_layers_color, hidden_dim_color
        self.encoder_dir, self.in_dim_dir = get_encoder(encoding_dir)
        # the extra colour grid is Seal-3D's change to torch-ngp (nerf/network.py:56,118)
        self.encoder_color, self.in_dim_color = get_encoder(encoding, desired_resolution=2048 * bound, log2_hashmap_size=log2_hashmap_size)
        self.color_net = nn.ModuleList([
            nn.Linear(self.in_dim_dir + geo_feat_dim + self.in_dim_color if l == 0 else hidden_dim_color,
                      3 if l == num_layers_color - 1 else hidden_dim_color, bias=False)
            for l in range(num_layers_color)])
        self.bg_net = None

    def _sigma_mlp(self, h):
        for l in range(self.num_layers):
            h = self.sigma_net[l](h)
            if l != self.num_layers - 1:
                h = F.relu(h, inplace=True)
        return h

    def _color_mlp(self, h):
        for l in range(self.num_layers_color):
            h = self.color_net[l](h)
            if l != self.num_layers_color - 1:
                h = F.relu(h, inplace=True)
        return torch.sigmoid(h)

    def forward(self, x, d):
        """nerf/network.py:99-128: x [N,3] in [-bound,bound], d [N,3] unit -> sigma [N], rgb [N,3]."""
        h = self._sigma_mlp(self.encoder(x, bound=self.bound))
        sigma = trunc_exp(h[..., 0])
        geo_feat = h[..., 1:]
        h = torch.cat([self.encoder_dir(d).to(geo_feat.dtype), geo_feat, self.encoder_color(x, bound=self.bound).to(geo_feat.dtype)], dim=-1)
        return sigma, self._color_mlp(h)

    def density(self, x):
        h = self._sigma_mlp(self.encoder(x, bound=self.bound))
        return {"sigma": trunc_exp(h[..., 0]), "geo_feat": h[..., 1:]}

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            x, d, geo_feat = x[mask], d[mask], geo_feat[mask]
        h = torch.cat([self.encoder_dir(d).to(geo_feat.dtype), geo_feat, self.encoder_color(x, bound=self.bound).to(geo_feat.dtype)], dim=-1)
        h = self._color_mlp(h)
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
            return rgbs
        return h

    def get_params(self, lr):
        """nerf/network.py:199-212 (group order = gradient-arena order)."""
        return [{"params": self.encoder.parameters(), "lr": lr}, {"params": self.sigma_net.parameters(), "lr": lr},
                {"params": self.encoder_color.parameters(), "lr": lr}, {"params": self.encoder_dir.parameters(), "lr": lr},
                {"params": self.color_net.parameters(), "lr": lr}]
