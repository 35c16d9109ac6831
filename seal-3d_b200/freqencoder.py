"""Host-side mirror of ``freqencoder/freq.py`` (``freq_encode`` / ``FreqEncoder``)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


class _freq_encoder(Function):
    """freq.py:15-51: inputs [B,D] -> [B, D + 2*D*degree] = (x, sin(2^f x), cos(2^f x), ...)."""

    @staticmethod
    def forward(ctx, inputs, degree, output_dim):
        inputs = inputs.cuda().contiguous().float()
        B, D = inputs.shape
        outputs = torch.empty(B, output_dim, dtype=torch.float32, device=inputs.device)
        _lib.call("s3d_freq_encode_forward", inputs, B, D, int(degree), int(output_dim), outputs)
        ctx.save_for_backward(inputs, outputs)
        ctx.dims = (B, D, int(degree), int(output_dim))
        return outputs

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous().float()
        inputs, outputs = ctx.saved_tensors
        B, D, degree, output_dim = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        _lib.call("s3d_freq_encode_backward", grad, outputs, B, D, degree, output_dim, grad_inputs)
        return grad_inputs, None, None


def freq_encode(inputs, degree, output_dim):
    return _freq_encoder.apply(inputs, degree, output_dim)


class FreqEncoder(nn.Module):
    """freq.py:56-76."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        prefix = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim)
        return freq_encode(inputs, self.degree, self.output_dim).reshape(prefix + [self.output_dim])
