"""`NeRFRenderer` with the contract of the reference's `lidarnerf/nerf/renderer.py:61-345` (subclass provides
`density(x) -> {"sigma","geo_feat"}` and `color(x, d, cal_lidar_color, mask, geo_feat)`; `render()` returns
`depth_lidar`, `image_lidar`, `weights_sum_lidar`) plus the occupancy-grid path the reference lacks:

  * `run`       - the reference's dense sampling (num_steps uniform + upsample_steps importance samples), restated;
                  pure torch, runs on CPU (BASELINE config 1) and is checked against tests/golden/ref_py_run.npz;
  * `run_cuda`  - `march_rays_train -> density/color -> composite_rays_train` on the sm_100a kernels op by op (B2
                  wrappers + torch autograd), with the density grid / bitfield / step-counter state of SURVEY.md
                  Appendix A kept on the module and refreshed every `grid_update_interval` (16) training calls (the
                  unmodified Trainer never calls update_extra_state);
  * `run_fused` - the same occupancy path as ONE autograd Function over the fused kernels (nerf/fused_render.py): what
                  the training engine launches, reachable from `Trainer.train_step` (nerf/utils.py:716-724).
`render()` picks the path.  The reference's CLI has no `cuda_ray` switch, so the default is
decided here: on a CUDA device, for `cal_lidar_color=True`, `render()` takes `run_fused` when the network maps onto the
fused kernels, else `run_cuda`; on the CPU (BASELINE config 1) or for the RGB branch it takes the reference's dense
`run`.  `LNB_CUDA_RAY=dense|ops|fused` (environment) or `render(..., cuda_ray="dense"|"ops"|"fused"|True|False)`
override.
LiDAR specifics kept: fixed near = min_near_lidar, far = 81 x near (renderer.py:129-138); absolute depth
(renderer.py:268; the training kernels' depth is relative to the march start and gets `+ t0 * weights_sum`,
SURVEY.md H2); no background blend.
"""
import math
import os

import torch
import torch.nn as nn


def sample_pdf(bins, weights, n_samples, det=False):
    """Inverse-CDF sampling of `n_samples` depths per ray from a piecewise-constant pdf (reference
    renderer.py:10-46).  bins [B,T], weights [B,T-1] -> [B,n_samples]."""
    w = weights + 1e-5
    cdf = torch.cumsum(w / w.sum(-1, keepdim=True), -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)                      # [B,T]
    shape = list(cdf.shape[:-1]) + [n_samples]
    if det:
        u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, n_samples, device=cdf.device).expand(shape)
    else:
        u = torch.rand(shape, device=cdf.device)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = (hi - 1).clamp(min=0)
    hi = hi.clamp(max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    return b_lo + (u - c_lo) / span * (b_hi - b_lo)


def _alpha_weights(sigmas, deltas, density_scale):
    alphas = 1 - torch.exp(-deltas * density_scale * sigmas)
    trans = torch.cumprod(torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], -1), -1)[..., :-1]
    return alphas * trans


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, density_scale=1, min_near=0.2, min_near_lidar=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.min_near_lidar = min_near_lidar
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius
        box = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        self.register_buffer("aabb_train", box)
        self.register_buffer("aabb_infer", box.clone())
        # occupancy state for run_cuda (persistent=False: checkpoints stay loadable by the reference)
        H3 = self.grid_size ** 3
        self.register_buffer("density_grid", torch.zeros(self.cascade, H3), persistent=False)
        self.register_buffer("density_bitfield", torch.full((self.cascade * H3 // 8,), 255, dtype=torch.uint8),
                             persistent=False)
        self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32), persistent=False)
        self.mean_density = 0.0
        self.mean_count = 0
        self.local_step = 0
        self.cuda_ray_calls = 0
        self.grid_update_interval = 16      # training calls between self-scheduled density-grid refreshes; 0 = the caller
                                            # manages the grid (update_extra_state / a user-written bitfield)
        self._fused = None                  # FusedLidarRender workspace (created on first use)

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    # ------------------------------------------------------------------------------------------------ bounds
    def _near_far(self, rays_o, rays_d, cal_lidar_color):
        n = rays_o.shape[0]
        if cal_lidar_color:
            nears = torch.full((n,), float(self.min_near_lidar), dtype=rays_o.dtype, device=rays_o.device)
            return nears, nears * 81.0
        from .. import raymarching
        aabb = self.aabb_train if self.training else self.aabb_infer
        return raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)

    # ------------------------------------------------------------------------------------------------ dense path
    def run(self, rays_o, rays_d, cal_lidar_color=False, num_steps=128, upsample_steps=128, bg_color=None,
            perturb=False, **kwargs):
        self.out_dim = self.out_lidar_color_dim if cal_lidar_color else self.out_color_dim
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = self._near_far(rays_o, rays_d, cal_lidar_color)
        nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)

        z = nears + (fars - nears) * torch.linspace(0.0, 1.0, num_steps, device=dev).unsqueeze(0)      # [N,T]
        step = (fars - nears) / num_steps
        if perturb:
            z = z + (torch.rand(z.shape, device=dev) - 0.5) * step

        def points(zv):
            p = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * zv.unsqueeze(-1)
            return torch.min(torch.max(p, aabb[:3]), aabb[3:])

        xyzs = points(z)
        dens = {k: v.view(N, num_steps, -1) for k, v in self.density(xyzs.reshape(-1, 3)).items()}

        if upsample_steps > 0:
            with torch.no_grad():
                d = torch.cat([z[..., 1:] - z[..., :-1], step * torch.ones_like(z[..., :1])], -1)
                w = _alpha_weights(dens["sigma"].squeeze(-1), d, self.density_scale)
                mid = z[..., :-1] + 0.5 * d[..., :-1]
                z_new = sample_pdf(mid, w[:, 1:-1], upsample_steps, det=not self.training).detach()
                xyz_new = points(z_new)
            dens_new = {k: v.view(N, upsample_steps, -1) for k, v in self.density(xyz_new.reshape(-1, 3)).items()}
            z, order = torch.sort(torch.cat([z, z_new], 1), dim=1)
            xyzs = torch.gather(torch.cat([xyzs, xyz_new], 1), 1, order.unsqueeze(-1).expand(-1, -1, 3))
            for k in dens:
                both = torch.cat([dens[k], dens_new[k]], 1)
                dens[k] = torch.gather(both, 1, order.unsqueeze(-1).expand_as(both))

        d = torch.cat([z[..., 1:] - z[..., :-1], step * torch.ones_like(z[..., :1])], -1)
        weights = _alpha_weights(dens["sigma"].squeeze(-1), d, self.density_scale)             # [N,T+t]
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        flat = {k: v.reshape(-1, v.shape[-1]) for k, v in dens.items()}
        mask = weights > 1e-4
        rgbs = self.color(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), cal_lidar_color=cal_lidar_color,
                          mask=mask.reshape(-1), **flat).view(N, -1, self.out_dim)
        weights_sum = weights.sum(-1)
        depth = (weights * z).sum(-1)
        image = (weights.unsqueeze(-1) * rgbs).sum(-2)
        if not cal_lidar_color:
            if self.bg_radius > 0:
                from .. import raymarching
                bg_color = self.background(raymarching.sph_from_ray(rays_o, rays_d, self.bg_radius), rays_d)
            elif bg_color is None:
                bg_color = 1
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        return {"depth_lidar": depth.view(*prefix), "image_lidar": image.view(*prefix, self.out_dim),
                "weights_sum_lidar": weights_sum}

    # ------------------------------------------------------------------------------------------------ occupancy path
    def run_cuda(self, rays_o, rays_d, cal_lidar_color=False, dt_gamma=0, perturb=False, force_all_rays=False,
                 max_steps=1024, T_thresh=1e-4, bg_color=None, **kwargs):
        from .. import raymarching as rm
        self.out_dim = self.out_lidar_color_dim if cal_lidar_color else self.out_color_dim
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        N = rays_o.shape[0]
        nears, fars = self._near_far(rays_o, rays_d, cal_lidar_color)

        if self.training:
            self._maybe_refresh_grid()
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            # the jitter is drawn HERE (not inside the wrapper) because the march start t0 = near + dt(near) * noise
            # (raymarching.cu:375) is needed below and the kernel does not return it
            noises = torch.rand(N, dtype=torch.float32, device=rays_o.device) if perturb else None
            xyzs, dirs, deltas, rays = rm.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield,
                                                           self.cascade, self.grid_size, nears, fars, counter,
                                                           self.mean_count, perturb, 128, force_all_rays, dt_gamma,
                                                           max_steps, noises)
            dens = self.density(xyzs)
            sigmas = dens.pop("sigma") * self.density_scale
            rgbs = self.color(xyzs, dirs, cal_lidar_color=cal_lidar_color, **dens).float()
            weights_sum, depth, image = rm.composite_rays_train(sigmas.float(), rgbs, deltas, rays, T_thresh, True)
            # the training kernel accumulates depth RELATIVE to the march start (raymarching.cu:640-646: `t` restarts at 0
            # on the first sample): add t0 * weights_sum so the LiDAR loss compares against absolute range
            t0 = nears
            if perturb:
                dt_min = 2 * math.sqrt(3) / max_steps
                dt_max = 2 * math.sqrt(3) * (1 << (self.cascade - 1)) / self.grid_size
                t0 = nears + torch.clamp(nears * dt_gamma, dt_min, dt_max) * noises
            depth = depth + t0 * weights_sum
        else:
            weights_sum = torch.zeros(N, dtype=torch.float32, device=rays_o.device)
            depth = torch.zeros_like(weights_sum)
            image = torch.zeros(N, 3, dtype=torch.float32, device=rays_o.device)
            alive = torch.arange(N, dtype=torch.int32, device=rays_o.device)
            rays_t = nears.clone()
            step = 0
            while step < max_steps and alive.numel() > 0:
                n_alive = alive.shape[0]
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, self.bound,
                                                   self.density_bitfield, self.cascade, self.grid_size, nears, fars,
                                                   128, perturb if step == 0 else False, dt_gamma, max_steps)
                dens = self.density(xyzs)
                sigmas = dens.pop("sigma") * self.density_scale
                rgbs = self.color(xyzs, dirs, cal_lidar_color=cal_lidar_color, **dens).float()
                if rgbs.shape[-1] < 3:                  # the inference kernel is 3-channel (raymarching.cu:1021-1023)
                    rgbs = torch.cat([rgbs, rgbs.new_zeros(rgbs.shape[0], 3 - rgbs.shape[-1])], -1)
                rm.composite_rays(n_alive, n_step, alive, rays_t, sigmas.float(), rgbs.contiguous(), deltas, weights_sum,
                                  depth, image, 1e-2)
                alive = alive[alive >= 0]
                step += n_step
            # (no `+ near * weights_sum` here: rays_t starts at `nears` and the inference kernel accumulates absolute t,
            # raymarching.cu:995-1025)
            image = image[:, : self.out_dim]
        if not cal_lidar_color:
            image = image + (1 - weights_sum).unsqueeze(-1) * (1 if bg_color is None else bg_color)
        return {"depth_lidar": depth.view(*prefix), "image_lidar": image.reshape(*prefix, self.out_dim),
                "weights_sum_lidar": weights_sum}

    def _maybe_refresh_grid(self):
        """Self-scheduled density-grid refresh of the occupancy paths (the unmodified Trainer never calls
        update_extra_state); `grid_update_interval = 0` turns it off."""
        k = int(self.grid_update_interval)
        if k > 0 and self.cuda_ray_calls % k == 0:
            self.update_extra_state()
        self.cuda_ray_calls += 1

    def fused_unsupported_reason(self):
        """None when `run_fused` can serve this network's LiDAR branch (subclasses that own the networks override)."""
        return "this renderer has no field network"

    def run_fused(self, rays_o, rays_d, cal_lidar_color=True, dt_gamma=0, perturb=False, max_steps=1024, T_thresh=1e-4,
                  **kwargs):
        """Occupancy march + fused field kernels + compositing as one autograd Function (nerf/fused_render.py)."""
        if not cal_lidar_color:
            raise RuntimeError("run_fused serves the LiDAR branch (cal_lidar_color=True) only")
        from .fused_render import FusedLidarRender
        if self._fused is None:
            self._fused = FusedLidarRender(self)
        self.out_dim = self.out_lidar_color_dim
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        if self.training:
            self._maybe_refresh_grid()
        weights_sum, depth, image = self._fused(rays_o, rays_d, perturb=perturb, dt_gamma=dt_gamma, max_steps=max_steps,
                                                T_thresh=T_thresh)
        return {"depth_lidar": depth.view(*prefix), "image_lidar": image.reshape(*prefix, self.out_dim),
                "weights_sum_lidar": weights_sum}

    def _pick_path(self, rays_o, cal_lidar_color, cuda_ray):
        if cuda_ray is None:
            cuda_ray = os.environ.get("LNB_CUDA_RAY", "auto")
        if cuda_ray is True:
            cuda_ray = "ops"
        if cuda_ray is False or cuda_ray in ("0", "dense", "off"):
            return self.run
        if cuda_ray == "ops":
            return self.run_cuda
        if cuda_ray == "fused":
            why = self.fused_unsupported_reason() if cal_lidar_color else "RGB branch"
            if why is not None:
                raise RuntimeError(f"cuda_ray='fused' requested but the fused kernels cannot serve this network: {why}")
            return self.run_fused
        if cuda_ray != "auto":
            raise ValueError(f"cuda_ray={cuda_ray!r}: choose from auto, dense, ops, fused")
        if not rays_o.is_cuda or not cal_lidar_color:
            return self.run
        return self.run_fused if self.fused_unsupported_reason() is None else self.run_cuda

    @torch.no_grad()
    def update_extra_state(self, decay=0.95):
        """Density-grid refresh: EMA-max of sigma at jittered cell centres, then packbits."""
        from .. import raymarching as rm
        H, dev = self.grid_size, self.density_grid.device
        idx = torch.arange(H ** 3, dtype=torch.int32, device=dev)
        coords = rm.morton3D_invert(idx).float()
        for cas in range(self.cascade):
            bound = min(2.0 ** cas, self.bound)
            half = bound / H
            xyz = (2 * coords / (H - 1) - 1) * (bound - half) + (torch.rand_like(coords) * 2 - 1) * half
            sig = torch.cat([self.density(chunk)["sigma"].reshape(-1).float() for chunk in xyz.split(2 ** 19)])
            self.density_grid[cas] = torch.maximum(self.density_grid[cas] * decay, sig * self.density_scale)
        self.mean_density = float(self.density_grid.clamp(min=0).mean().item())
        rm.packbits(self.density_grid, min(self.mean_density, self.density_thresh), self.density_bitfield)
        used = min(self.local_step, 16)
        if used > 0:
            self.mean_count = int(self.step_counter[:used, 0].sum().item() / used)
        self.local_step = 0

    # ------------------------------------------------------------------------------------------------ entry point
    def render(self, rays_o, rays_d, cal_lidar_color=False, staged=False, max_ray_batch=4096, cuda_ray=None,
               **kwargs):
        fn = self._pick_path(rays_o, cal_lidar_color, cuda_ray)
        if not staged:
            return fn(rays_o, rays_d, cal_lidar_color=cal_lidar_color, **kwargs)
        B, N = rays_o.shape[:2]
        out_dim = self.out_lidar_color_dim if cal_lidar_color else self.out_color_dim
        depth = torch.empty((B, N), device=rays_o.device)
        image = torch.empty((B, N, out_dim), device=rays_o.device)
        for b in range(B):
            for head in range(0, N, max_ray_batch):
                tail = min(head + max_ray_batch, N)
                part = fn(rays_o[b:b + 1, head:tail], rays_d[b:b + 1, head:tail], cal_lidar_color=cal_lidar_color,
                          **kwargs)
                depth[b:b + 1, head:tail] = part["depth_lidar"]
                image[b:b + 1, head:tail] = part["image_lidar"]
        return {"depth_lidar": depth, "image_lidar": image}
