"""`trunc_exp` of the reference (`lidarnerf/activation.py:6-20`): exp in fp32, gradient with the exponent
clamped to [-15, 15]."""
import torch
from torch.autograd import Function


class _TruncExp(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply
