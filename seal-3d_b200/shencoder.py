"""Host-side mirror of ``shencoder/sphere_harmonics.py`` (``sh_encode`` / ``SHEncoder``)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


class _sh_encoder(Function):
    """sphere_harmonics.py:14-52: inputs [B,3] float32 -> [B, degree^2]; always float32."""

    @staticmethod
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous().float()
        B, D = inputs.shape
        C2 = degree ** 2
        outputs = torch.empty(B, C2, dtype=torch.float32, device=inputs.device)
        dy_dx = torch.empty(B, D * C2, dtype=torch.float32, device=inputs.device) if calc_grad_inputs else None
        _lib.check_cuda(inputs)
        _lib.call("s3d_sh_encode_forward", inputs, outputs, B, D, int(degree), dy_dx)
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims = (B, D, int(degree))
        return outputs

    @staticmethod
    def backward(ctx, grad):
        inputs, dy_dx = ctx.saved_tensors
        if dy_dx is None:
            return None, None, None
        B, D, degree = ctx.dims
        grad = grad.contiguous().float()
        grad_inputs = torch.zeros_like(inputs)
        _lib.call("s3d_sh_encode_backward", grad, inputs, B, D, degree, dy_dx, grad_inputs)
        return grad_inputs, None, None


def sh_encode(inputs, degree, calc_grad_inputs=False):
    return _sh_encoder.apply(inputs, degree, calc_grad_inputs)


class SHEncoder(nn.Module):
    """sphere_harmonics.py:59-86."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        out = sh_encode(inputs, self.degree, inputs.requires_grad)
        return out.reshape(prefix + [self.output_dim])
