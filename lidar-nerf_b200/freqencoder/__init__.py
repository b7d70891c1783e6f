"""Boundary B2: `freq_encode` / `FreqEncoder` of the reference's `lidarnerf/freqencoder/freq.py` (:12-77)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from ..backend import _freqencoder as _backend


class _FreqEncode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        inputs = (inputs if inputs.is_cuda else inputs.cuda()).contiguous()
        B, D = inputs.shape
        outputs = inputs.new_empty(B, output_dim)
        _backend.freq_encode_forward(inputs, B, D, degree, output_dim, outputs)
        ctx.save_for_backward(outputs)
        ctx.cfg = (B, D, degree, output_dim)
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        (outputs,) = ctx.saved_tensors
        B, D, degree, output_dim = ctx.cfg
        grad_inputs = outputs.new_zeros(B, D)
        _backend.freq_encode_backward(grad.contiguous(), outputs, B, D, degree, output_dim, grad_inputs)
        return grad_inputs, None, None


freq_encode = _FreqEncode.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        lead = list(inputs.shape[:-1])
        out = freq_encode(inputs.reshape(-1, self.input_dim), self.degree, self.output_dim)
        return out.reshape(lead + [self.output_dim])


__all__ = ["freq_encode", "FreqEncoder"]
