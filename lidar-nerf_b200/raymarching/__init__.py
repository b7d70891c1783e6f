"""Boundary B2 for ray marching / compositing: the public functions of the reference's
`lidarnerf/raymarching/raymarching.py` (near_far_from_aabb :48, sph_from_ray :82, morton3D :108,
morton3D_invert :133, packbits :164, march_rays_train :289, composite_rays_train :360, march_rays :460,
composite_rays :510) with the same argument meaning, defaults, return values and error behaviour, running
on the sm_100a kernels of liblnb200.so.

Extensions (opt-in, never change the defaults): `composite_rays_train(..., depth_grad=True)` propagates the
depth gradient the reference drops (raymarching.py:329-330), and rgbs may have 1-4 channels
(the reference hard-codes 3; the LiDAR head emits 2).
"""
import torch
from torch.autograd import Function

from ..backend import _raymarching as _backend

_fwd32 = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = torch.amp.custom_bwd(device_type="cuda")


def _rays_2d(t):
    t = t if t.is_cuda else t.cuda()
    return t.contiguous().view(-1, 3)


class _NearFar(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o, rays_d = _rays_2d(rays_o), _rays_2d(rays_d)
        n = rays_o.shape[0]
        nears = rays_o.new_empty(n)
        fars = rays_o.new_empty(n)
        _backend.near_far_from_aabb(rays_o, rays_d, aabb.contiguous(), n, min_near, nears, fars)
        return nears, fars


class _SphFromRay(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        rays_o, rays_d = _rays_2d(rays_o), _rays_2d(rays_d)
        coords = rays_o.new_empty(rays_o.shape[0], 2)
        _backend.sph_from_ray(rays_o, rays_d, radius, rays_o.shape[0], coords)
        return coords


class _Morton(Function):
    @staticmethod
    def forward(ctx, coords):
        coords = (coords if coords.is_cuda else coords.cuda()).int().contiguous()
        out = torch.empty(coords.shape[0], dtype=torch.int32, device=coords.device)
        _backend.morton3D(coords, coords.shape[0], out)
        return out


class _MortonInvert(Function):
    @staticmethod
    def forward(ctx, indices):
        indices = (indices if indices.is_cuda else indices.cuda()).int().contiguous()
        out = torch.empty(indices.shape[0], 3, dtype=torch.int32, device=indices.device)
        _backend.morton3D_invert(indices, indices.shape[0], out)
        return out


class _Packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        grid = (grid if grid.is_cuda else grid.cuda()).contiguous()
        n = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(n, dtype=torch.uint8, device=grid.device)
        _backend.packbits(grid, n, thresh, bitfield)
        return bitfield


class _MarchTrain(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, noises=None):
        # `noises` (extension, optional, trailing): the per-ray jitter in [0,1) the caller drew itself because it needs
        # the march start t0 = near + dt * noise afterwards; None = drawn here like the reference (raymarching.py:250-253)
        rays_o, rays_d = _rays_2d(rays_o), _rays_2d(rays_d)
        bitfield = (density_bitfield if density_bitfield.is_cuda else density_bitfield.cuda()).contiguous()
        dev, dt = rays_o.device, rays_o.dtype
        n = rays_o.shape[0]

        # sample budget: worst case until a running mean of real counts is known (raymarching.py:226-233)
        budget = n * max_steps
        use_mean = (not force_all_rays) and mean_count > 0
        if use_mean:
            if align > 0:
                mean_count += align - mean_count % align
            budget = mean_count

        xyzs = torch.zeros(budget, 3, dtype=dt, device=dev)
        dirs = torch.zeros(budget, 3, dtype=dt, device=dev)
        deltas = torch.zeros(budget, 2, dtype=dt, device=dev)
        rays = torch.empty(n, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        if noises is None:
            noises = torch.rand(n, dtype=dt, device=dev) if perturb else torch.zeros(n, dtype=dt, device=dev)
        noises = noises.to(dt).contiguous()

        _backend.march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, n, C, H, budget,
                                  nears.contiguous(), fars.contiguous(), xyzs, dirs, deltas, rays, step_counter,
                                  noises)

        if not use_mean:  # trim to the produced count (one D2H sync, as in the reference's first epochs)
            m = int(step_counter[0].item())
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


class _CompositeTrain(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4, depth_grad=False):
        sigmas, rgbs, deltas = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous()
        m, n, ch = sigmas.shape[0], rays.shape[0], rgbs.shape[-1]
        weights_sum = sigmas.new_empty(n)
        depth = sigmas.new_empty(n)
        image = sigmas.new_empty(n, ch)
        if ch == 3:
            _backend.composite_rays_train_forward(sigmas, rgbs, deltas, rays, m, n, T_thresh, weights_sum, depth, image)
        else:
            _backend.composite_rays_train_forward_ex(sigmas, rgbs, deltas, rays, m, n, T_thresh, ch, weights_sum,
                                                     depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.cfg = (m, n, T_thresh, ch, bool(depth_grad))
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, g_ws, g_depth, g_image):
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        m, n, T_thresh, ch, depth_grad = ctx.cfg
        g_ws, g_image = g_ws.contiguous(), g_image.contiguous()
        g_sigmas, g_rgbs = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        if ch == 3 and not depth_grad:
            _backend.composite_rays_train_backward(g_ws, g_image, sigmas, rgbs, deltas, rays, weights_sum, image, m, n,
                                                   T_thresh, g_sigmas, g_rgbs)
        else:
            gd = g_depth.contiguous() if depth_grad else None
            _backend.composite_rays_train_backward_ex(g_ws, gd, g_image, sigmas, rgbs, deltas, rays, weights_sum,
                                                      depth if depth_grad else None, image, m, n, T_thresh, ch,
                                                      g_sigmas, g_rgbs)
        return g_sigmas, g_rgbs, None, None, None, None


class _MarchInfer(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        rays_o, rays_d = _rays_2d(rays_o), _rays_2d(rays_d)
        dev, dt = rays_o.device, rays_o.dtype
        m = n_alive * n_step
        if align > 0:
            m += align - (m % align)
        xyzs = torch.zeros(m, 3, dtype=dt, device=dev)
        dirs = torch.zeros(m, 3, dtype=dt, device=dev)
        deltas = torch.zeros(m, 2, dtype=dt, device=dev)
        noises = torch.rand(n_alive, dtype=dt, device=dev) if perturb else torch.zeros(n_alive, dtype=dt, device=dev)
        _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                            density_bitfield.contiguous(), near, far, xyzs, dirs, deltas, noises)
        return xyzs, dirs, deltas


class _CompositeInfer(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                T_thresh=1e-2):
        _backend.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas.contiguous(), rgbs.contiguous(),
                                deltas, weights_sum, depth, image)
        return tuple()


near_far_from_aabb = _NearFar.apply
sph_from_ray = _SphFromRay.apply
morton3D = _Morton.apply
morton3D_invert = _MortonInvert.apply
packbits = _Packbits.apply
march_rays_train = _MarchTrain.apply
composite_rays_train = _CompositeTrain.apply
march_rays = _MarchInfer.apply
composite_rays = _CompositeInfer.apply

__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train", "march_rays", "composite_rays"]
