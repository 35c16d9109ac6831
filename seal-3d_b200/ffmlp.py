"""Host-side mirror of ``ffmlp/ffmlp.py`` (``ffmlp_forward`` / ``FFMLP``) on the tcgen05 kernels.

The reference wrapper cannot even be imported here (``from turtle import backward, forward`` needs
tkinter, ffmlp.py:2); semantics follow ffmlp.py:15-168: fp16 tensors, flat weight vector
``[hidden*in | hidden*hidden*(num_layers-1) | padded_out*hidden]`` (row-major [out,in] blocks),
U(+-sqrt(3/hidden)) init with seed 42, output padded to 16, batch padded to 128 by the caller in the
reference (not needed here: the kernels mask the tail tile).
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib


def convert_activation(act):
    return {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5}.get(act, 6)


class _ffmlp_forward(Function):
    @staticmethod
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                inference=False, calc_grad_inputs=False):
        inputs = inputs.contiguous().half()
        weights = weights.contiguous().half()
        B = inputs.shape[0]
        outputs = torch.empty(B, output_dim, device=inputs.device, dtype=torch.float16)
        _lib.check_cuda(inputs, weights)
        if not inference:
            forward_buffer = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=torch.float16)
            _lib.call("s3d_ffmlp_forward", inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                      output_activation, forward_buffer, outputs)
            ctx.save_for_backward(inputs, weights, outputs, forward_buffer)
            ctx.dims = (input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs)
        else:
            _lib.call("s3d_ffmlp_inference", inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                      output_activation, None, outputs)
        return outputs

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous().half()
        B = grad.shape[0]
        inputs, weights, outputs, forward_buffer = ctx.saved_tensors
        input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs = ctx.dims
        grad_inputs = torch.zeros_like(inputs) if calc_grad_inputs else None
        grad_weights = torch.empty_like(weights)
        backward_buffer = torch.empty(num_layers, B, hidden_dim, device=grad.device, dtype=torch.float16)
        _lib.call("s3d_ffmlp_backward", grad, inputs, weights, forward_buffer, B, input_dim, output_dim, hidden_dim, num_layers,
                  activation, output_activation, int(bool(calc_grad_inputs)), backward_buffer, grad_inputs, grad_weights)
        return grad_inputs, grad_weights, None, None, None, None, None, None, None, None


def ffmlp_forward(*args):
    return _ffmlp_forward.apply(*args)


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation="relu"):
        super().__init__()
        self.input_dim, self.output_dim, self.hidden_dim, self.num_layers = input_dim, output_dim, hidden_dim, num_layers
        self.activation = convert_activation(activation)
        self.output_activation = convert_activation("none")
        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"
        self.padded_output_dim = int(math.ceil(output_dim / 16)) * 16
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        _lib.call_nostream("s3d_allocate_splitk", self.num_layers + 1)

    def cleanup(self):
        _lib.call_nostream("s3d_free_splitk")

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def reset_parameters(self):
        torch.manual_seed(42)
        std = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-std, std)

    def forward(self, inputs):
        B = inputs.shape[0]
        outputs = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers,
                                self.activation, self.output_activation, not self.training, inputs.requires_grad)
        if self.padded_output_dim != self.output_dim:
            outputs = outputs[:B, :self.output_dim]
        return outputs
