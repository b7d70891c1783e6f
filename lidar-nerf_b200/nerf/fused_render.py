"""The fused LiDAR-field step behind the reference's call path.

`Trainer.train_step` of the reference calls `self.model.render(rays_o, rays_d, cal_lidar_color=True, staged=False,
perturb=True, **vars(opt))`, builds the LiDAR loss from `depth_lidar` / `image_lidar` in torch, and drives
`loss.backward(); optimizer.step()` (nerf/utils.py:697-734, 1206-1226).  `FusedLidarRender` serves that call with ONE
`torch.autograd.Function`:

  forward   occupancy march -> hash-grid gather -> fused field kernel (density MLP + LiDAR head, tcgen05) -> compositing
  backward  compositing backward (with the depth gradient) -> LiDAR-head / density-MLP backward (tcgen05, weight
            gradients accumulated in tensor memory) -> hash-grid scatter, written straight into the `.grad` of
            `encoder.embeddings`, `sigma_net.weights` and `lidar_color_net.weights`

so the unmodified Trainer (GradScaler, torch.optim.Adam, EMA, checkpoints) trains the sm_100a kernels.  The kernels and
their launch order are the training engine's (nerf/engine.py); the engine object here is only a WORKSPACE - it owns the
sample / activation buffers and an fp16 shadow of the parameters that is refreshed from the module's fp32 parameters
every call, the way the reference casts its tables per forward (grid.py:45-46, ffmlp.py:33-36).
"""
import math

import torch
from torch.autograd import Function

from .config import FieldConfig


class _FusedLidarField(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, embeddings, w_sigma, w_head, ws, perturb):
        eng = ws.engine(rays_o.shape[0], rays_o.device)
        n = rays_o.shape[0]
        eng.cfg.perturb = bool(perturb)
        eng.load_params(embeddings, w_sigma, w_head)
        eng.rays_o.copy_(rays_o)
        eng.rays_d.copy_(rays_d)
        eng._fb_march()
        eng._fwd_field()
        eng.composite_forward()
        ws.after_march(eng)
        # the kernels measure depth from the (jittered) march start t0 (raymarching.cu:375,640-646): the LiDAR loss
        # compares against absolute range, so add t0 * weights_sum back
        depth = torch.addcmul(eng.depth, eng.t0, eng.ws)
        ctx.ws, ctx.n = ws, n
        ctx.key = ws.stamp(eng)
        ctx.set_materialize_grads(True)
        return eng.ws.clone(), depth, eng.image.clone()

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_ws, g_depth, g_image):
        ws = ctx.ws
        eng = ws.engine(ctx.n, g_ws.device)
        if ws.stamp(eng) != ctx.key:
            raise RuntimeError("FusedLidarRender: backward() of a forward whose workspace has been reused by a later "
                               "render() call - call backward before rendering the next training batch")
        # d(depth_abs)/d(ws) = t0
        torch.addcmul(g_ws.float(), g_depth.float(), eng.t0, out=eng.g_ws)
        eng.g_depth.copy_(g_depth)
        eng.g_image.copy_(g_image)
        G = torch.empty(eng.ex.n_padded, dtype=torch.float32, device=g_ws.device)   # zeroed by the engine (late_grad_zero)
        eng.set_grad_buffer(G)
        eng.composite_backward()
        eng._bwd_field()
        c = eng.cfg
        g_emb = eng.g_table.view(eng.n_rows, c.level_dim)
        return None, None, g_emb, eng.g_sigma_w, eng.g_head_w, None, None


class FusedLidarRender:
    """Workspace + dispatcher owned by a NeRFNetwork: `__call__` returns (weights_sum, depth, image) for one ray batch."""

    def __init__(self, net):
        self.net = net
        self._engines = {}
        self._serial = 0

    # ---- what the fused kernels implement -------------------------------------------------------------------------
    @staticmethod
    def supported(net):
        """None when the network's LiDAR branch maps onto the fused kernels, else the reason it does not."""
        from ..gridencoder import GridEncoder
        from ..freqencoder import FreqEncoder
        if not getattr(net, "use_ffmlp", False):
            return "MLPs are nn.Linear stacks (use_ffmlp=False)"
        enc = net.encoder
        if not isinstance(enc, GridEncoder):
            return "position encoder is not a GridEncoder"
        if enc.input_dim != 3 or enc.gridtype != "hash" or enc.align_corners or enc.interpolation != "linear":
            return "GridEncoder variant (needs 3-D hash grid, linear interpolation, align_corners=False)"
        if enc.embeddings.shape[1] != 2:
            return "level_dim != 2"
        if not isinstance(net.encoder_lidar_dir, FreqEncoder):
            return "LiDAR direction encoder is not a FreqEncoder"
        if net.geo_feat_dim != 15 or net.out_lidar_color_dim != 2:
            return "geo_feat_dim != 15 or out_lidar_color_dim != 2"
        if net.sigma_net.hidden_dim != 64 or net.lidar_color_net.hidden_dim != 64:
            return "hidden_dim != 64"
        from .._lib import lib, u32
        rc = lib.lnb_field_supported(u32(enc.output_dim), u32(net.sigma_net.num_layers), u32(net.lidar_color_net.input_dim),
                                     u32(net.lidar_color_net.num_layers), u32(net.encoder_lidar_dir.degree), u32(64))
        return None if rc == 0 else f"lnb_field_supported status {rc}"

    def config(self, **over):
        net, enc = self.net, self.net.encoder
        top = enc.base_resolution * enc.per_level_scale ** (enc.num_levels - 1)
        c = FieldConfig(bound=float(net.bound), grid_size=net.grid_size, min_near_lidar=float(net.min_near_lidar),
                        density_scale=float(net.density_scale), density_thresh=float(net.density_thresh),
                        num_levels=enc.num_levels, level_dim=enc.level_dim, base_resolution=enc.base_resolution,
                        desired_resolution=int(round(top)), log2_hashmap_size=enc.log2_hashmap_size,
                        hidden_dim=64, sigma_layers=net.sigma_net.num_layers, head_layers=net.lidar_color_net.num_layers,
                        freq_degree=net.encoder_lidar_dir.degree, geo_feat_dim=net.geo_feat_dim,
                        fused_composite=False, compact_backward=False, late_grad_zero=True, grid_update_interval=0)
        for k, v in over.items():
            setattr(c, k, v)
        return c

    def engine(self, n_rays, device):
        from .engine import LidarFieldEngine
        key = (int(n_rays), str(device))
        eng = self._engines.get(key)
        if eng is None:
            c = self.config(**self._march)
            near, far = c.min_near_lidar, c.min_near_lidar * c.far_factor
            dt_min = 2 * math.sqrt(3) / c.max_steps
            per_ray = min(c.max_steps, int((far - near) / dt_min) + 2)      # worst case: every step emits a sample
            eng = LidarFieldEngine(c, n_rays, device=device, sample_budget=n_rays * per_ray, external_params=True)
            # the level scale must be EXACTLY the module's (the engine derives it from desired_resolution)
            eng.S = float(math.log2(self.net.encoder.per_level_scale))
            eng.per_level_scale = float(self.net.encoder.per_level_scale)
            assert eng.n_rows == self.net.encoder.embeddings.shape[0], "level table sizing differs from GridEncoder"
            self._engines[key] = eng
            if len(self._engines) > 4:                                       # ray-count churn (staged eval tails): keep 4
                self._engines.pop(next(iter(self._engines)))
        # occupancy state lives on the module (shared by every workspace and by run_cuda)
        eng.bitfield = self.net.density_bitfield
        return eng

    def stamp(self, eng):
        return (id(eng), eng._serial if hasattr(eng, "_serial") else 0)

    def after_march(self, eng):
        self._serial += 1
        eng._serial = self._serial

    _march = {}

    def __call__(self, rays_o, rays_d, perturb=False, dt_gamma=0.0, max_steps=1024, T_thresh=1e-4):
        net = self.net
        march = dict(dt_gamma=float(dt_gamma), max_steps=int(max_steps), T_thresh=float(T_thresh))
        if march != self._march:                 # march law changed: the workspaces were sized / configured for the old one
            self._march = march
            self._engines.clear()
        return _FusedLidarField.apply(rays_o, rays_d, net.encoder.embeddings, net.sigma_net.weights,
                                      net.lidar_color_net.weights, self, bool(perturb))
