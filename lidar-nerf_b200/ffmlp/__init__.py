"""Boundary B2: `ffmlp_forward` / `FFMLP` of the reference's `lidarnerf/ffmlp/ffmlp.py` (:14-167, :187-283).

Same constructor constraints and flat fp16 weight layout ([hidden*in | (L-1)*hidden^2 | 16*hidden]), same
initialisation (seed 42, U(+-sqrt(3/hidden)), ffmlp.py:242-245) and the same padding rule (pads
128 - B % 128 rows even when B % 128 == 0, ffmlp.py:257-262).  The kernels run on the tcgen05 tensor cores with
fp32 accumulation; this build supports hidden_dim 64 natively and 16 / 32 through zero-padded weights on the same
kernels (backend._widen_index), ReLU, input_dim <= 128, and raises RuntimeError otherwise (the reference also
supports 128/256 and other activations).
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function

from ..backend import _ffmlp as _backend

_ACTIVATIONS = {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5}


def convert_activation(act):
    return _ACTIVATIONS.get(act, 6)


class _FFMLPFunction(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation,
                inference=False, calc_grad_inputs=False):
        inputs, weights = inputs.contiguous(), weights.contiguous()
        B = inputs.shape[0]
        outputs = torch.empty(B, output_dim, device=inputs.device, dtype=inputs.dtype)
        if inference:
            _backend.ffmlp_inference(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                                     output_activation, None, outputs)
            return outputs
        forward_buffer = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
        _backend.ffmlp_forward(inputs, weights, B, input_dim, output_dim, hidden_dim, num_layers, activation,
                               output_activation, forward_buffer, outputs)
        ctx.save_for_backward(inputs, weights, forward_buffer)
        ctx.cfg = (input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs)
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, weights, forward_buffer = ctx.saved_tensors
        input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs = ctx.cfg
        grad = grad.contiguous()
        B = grad.shape[0]
        grad_inputs = torch.empty_like(inputs) if calc_grad_inputs else None
        grad_weights = torch.empty_like(weights)
        # the reference's [num_layers, B, hidden] backward_buffer never leaves the SM here -> not allocated
        _backend.ffmlp_backward(grad, inputs, weights, forward_buffer, B, input_dim, output_dim, hidden_dim,
                                num_layers, activation, output_activation, calc_grad_inputs, None, grad_inputs,
                                grad_weights)
        return grad_inputs, grad_weights, None, None, None, None, None, None, None, None


ffmlp_forward = _FFMLPFunction.apply


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation="relu"):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.activation = convert_activation(activation)
        self.output_activation = convert_activation("none")
        self.tensorcore_width = 16

        assert hidden_dim in [16, 32, 64, 128, 256], \
            f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"

        self.padded_output_dim = int(math.ceil(output_dim / 16)) * 16
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()
        _backend.allocate_splitk(self.num_layers + 1)

    def cleanup(self):
        _backend.free_splitk()

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def reset_parameters(self):
        torch.manual_seed(42)
        bound = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-bound, bound)

    def forward(self, inputs):
        B, C = inputs.shape
        pad = 128 - (B % 128)
        if pad > 0:
            inputs = torch.cat([inputs, torch.zeros(pad, C, dtype=inputs.dtype, device=inputs.device)], dim=0)
        outputs = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim,
                                self.num_layers, self.activation, self.output_activation, not self.training,
                                inputs.requires_grad)
        if B != outputs.shape[0] or self.padded_output_dim != self.output_dim:
            outputs = outputs[:B, : self.output_dim]
        return outputs


__all__ = ["ffmlp_forward", "FFMLP", "convert_activation"]
