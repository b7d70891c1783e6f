"""Boundary B2 for the multiresolution hash/tiled grid encoder: `grid_encode` and `GridEncoder` of the
reference's `lidarnerf/gridencoder/grid.py` (:24-138, :141-235) - same constructor arguments, attributes
(`embeddings`, `offsets`, `output_dim`, ...), level sizing (:179-192), init (:202-204), autocast rule
(:54-57) and forward semantics (inputs in [-bound, bound], :209-235) - on the sm_100a kernels.

Difference in mechanism only: the kernels read/write the [B, L*C] layout directly, so the reference's
`permute(1,0,2).reshape` after forward and `.permute(1,0,2).contiguous()` before backward (grid.py:87,104)
disappear; values are identical.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from ..backend import _gridencoder as _backend

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}
_LAYOUT_BLC = 1


class _GridEncode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)

        # only the table goes to half under autocast, and only for even C (grid.py:54-57)
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        embeddings = embeddings.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype) if calc_grad_inputs else None
        _backend.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                     align_corners, interpolation, layout=_LAYOUT_BLC)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.cfg = (B, D, C, L, S, H, gridtype, interpolation, align_corners)
        return outputs

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, align_corners = ctx.cfg
        grad = grad.to(embeddings.dtype).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
        _backend.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                                      grad_inputs, gridtype, align_corners, interpolation, layout=_LAYOUT_BLC)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _GridEncode.apply


def level_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size, align_corners):
    """Rows per level, rounded up to 8 and capped at 2^log2_hashmap_size (grid.py:179-192)."""
    cap = 2 ** log2_hashmap_size
    sizes = []
    for lvl in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** lvl))
        rows = min(cap, (res if align_corners else res + 1) ** input_dim)
        sizes.append(int(np.ceil(rows / 8) * 8))
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype="hash", align_corners=False,
                 interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:  # overrides per_level_scale (grid.py:158-161)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners
        self.max_params = 2 ** log2_hashmap_size

        offsets = level_offsets(input_dim, num_levels, base_resolution, per_level_scale, log2_hashmap_size,
                                align_corners)
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        top = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {top} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} align_corners={self.align_corners} "
                f"interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)  # [-bound, bound] -> [0, 1]
        lead = list(inputs.shape[:-1])
        flat = inputs.view(-1, self.input_dim)
        out = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          flat.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return out.view(lead + [self.output_dim])

    @torch.no_grad()
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """grid.py:238-277: add the total-variation gradient at `inputs` (in [-bound, bound]; B random points when None)
        to `embeddings.grad`.  Call after loss.backward() and before optimizer.step()."""
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        L = self.offsets.shape[0] - 1
        _backend.grad_total_variation(inputs.to(self.embeddings.dtype).contiguous(), self.embeddings, self.embeddings.grad,
                                      self.offsets, weight, B, self.input_dim, self.embeddings.shape[1], L,
                                      float(np.log2(self.per_level_scale)), self.base_resolution, self.gridtype_id,
                                      self.align_corners)


__all__ = ["grid_encode", "GridEncoder", "level_offsets"]
