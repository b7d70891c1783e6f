"""Fused LiDAR-field training engine: one optimiser step of the reference's hot path
(march -> hash-grid -> density MLP -> LiDAR head -> composite -> loss -> backward -> Adam) as 9 launches of
liblnb200.so kernels, captured in a CUDA graph, with no autograd, no host sync and no allocation inside the step.

What it replaces in the reference (SURVEY.md sections 3.1-3.2): `Trainer.train_step` (nerf/utils.py:697-734) +
`NeRFRenderer.run` (nerf/renderer.py:99-298) + `NeRFNetwork.density/color` (nerf/network.py:162-237) +
`torch.optim.Adam`/GradScaler (main_lidarnerf.py:389-391, nerf/utils.py:1221-1223), in occupancy-march mode
(the `run_cuda` glue the reference lacks).

Parameters live in ONE flat fp32 vector [hash table | density-MLP weights | LiDAR-head weights] with a flat
fp16 shadow the kernels read, so Adam and the data-parallel gradient exchange are single passes over one buffer.

The step, in launch order:
  march (+ tail zeroing)  ||  per-ray direction terms        k_march_train, k_ray_dir_terms
  hash-grid gather                                            k_grid_fwd
  density MLP -> sigma/geo -> LiDAR head                      k_field_fwd            (tcgen05 + TMEM)
  composite fwd + LiDAR loss + composite bwd + live-row list  k_lidar_composite_step
  head backward / density-MLP backward on the live rows       k_mlp_bwd<head>, k_mlp_bwd   (tcgen05 + TMEM)
  gradient-table memset, hash-grid scatter on the live rows   k_grid_bwd
  Adam (one rank)  |  peer-memory reduce-scatter + Adam + all-gather (data parallel)   k_adam | k_dp_adam_exchange
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from ..backend import (_raymarching as rm, _ffmlp as ff, adam_step)
from .._lib import lib, check, u32, f32, i32, vp, sz, launch_count
from ..gridencoder import level_offsets
from . import dp
from .config import FieldConfig   # noqa: F401  (re-exported: `from ...engine import FieldConfig` keeps working)


def _ck(status, what):
    check(status, what)


class LidarFieldEngine:
    def __init__(self, cfg: FieldConfig, n_rays: int, device="cuda:0", sample_budget: int = None,
                 external_params: bool = False):
        """external_params=True: the engine is a WORKSPACE for parameters owned by someone else (the nn.Module path,
        nerf/fused_render.py): only the fp16 shadow `Ph` exists (filled by `load_params`), no fp32 master / Adam moments /
        gradient vector (the caller hands a gradient buffer to `set_grad_buffer` before the backward kernels run)."""
        self.cfg = cfg
        self.external_params = bool(external_params)
        self._sfx = "_bf16" if cfg.mlp_dtype == "bf16" else ""
        self.dev = torch.device(device)
        self.N = int(n_rays)
        c = cfg
        dev = self.dev
        gen = torch.Generator(device="cpu").manual_seed(c.seed)

        # ---- parameters -------------------------------------------------------------------------------------
        pls = float(np.exp2(np.log2(c.desired_resolution / c.base_resolution) / (c.num_levels - 1)))
        self.per_level_scale = pls
        self.S = float(np.log2(pls))
        offs = level_offsets(3, c.num_levels, c.base_resolution, pls, c.log2_hashmap_size, False)
        self.offsets = torch.from_numpy(offs).to(dev)
        self.n_rows = int(offs[-1])
        n_table = self.n_rows * c.level_dim
        self.enc_dim = c.num_levels * c.level_dim
        n_sigma = c.hidden_dim * (self.enc_dim + c.hidden_dim * (c.sigma_layers - 1) + 16)
        n_head = c.hidden_dim * (c.head_in_dim + c.hidden_dim * (c.head_layers - 1) + 16)
        self.n_table, self.n_sigma, self.n_head = n_table, n_sigma, n_head
        n = n_table + n_sigma + n_head
        self.n_params = n
        self.ex = dp.ShardedExchange(n, local_only=self.external_params)   # data-parallel layout of the flat vectors
        npad = self.ex.n_padded
        self._peer = None
        if self.external_params:
            P = None
            self.Ph = torch.zeros(npad, dtype=torch.float16, device=dev)
            self.G = None
        else:
            P = torch.zeros(npad, dtype=torch.float32)
            P[:n_table].uniform_(-1e-4, 1e-4, generator=gen)                                   # grid.py:202-204
            bound_w = math.sqrt(3 / c.hidden_dim)                                              # ffmlp.py:242-245
            P[n_table:n].uniform_(-bound_w, bound_w, generator=gen)
            self.G = torch.zeros(npad, dtype=torch.float32, device=dev)
            self.Ph = P.to(dev).to(torch.float16)
        if self.ex.world > 1 and c.fused_exchange and not self.external_params:
            # gradient and fp16 shadow in symmetric (peer-mapped) memory: the exchange becomes ONE kernel that reads every
            # rank's gradient shard and writes every rank's shadow over NVLink (lnb_dp_adam_exchange); NCCL otherwise
            try:
                import torch.distributed as dist
                import torch.distributed._symmetric_memory as symm
                Gs = symm.empty(npad, dtype=torch.float32, device=dev)
                hG = symm.rendezvous(Gs, dist.group.WORLD)
                Ps = symm.empty(npad, dtype=torch.float16, device=dev)
                hP = symm.rendezvous(Ps, dist.group.WORLD)
                Gs.zero_()
                Ps.copy_(self.Ph)
                ptr_t = C.c_void_p * self.ex.world
                self._peer = dict(hG=hG, hP=hP, g=ptr_t(*[int(x) for x in hG.buffer_ptrs]),
                                  h=ptr_t(*[int(x) for x in hP.buffer_ptrs]), mc=None)
                mc_g, mc_h = int(getattr(hG, "multicast_ptr", 0) or 0), int(getattr(hP, "multicast_ptr", 0) or 0)
                if c.multicast_exchange and mc_g and mc_h:
                    self._peer["mc"] = (mc_g, mc_h)          # NVLS: in-switch reduce-scatter + multicast all-gather
                self.G, self.Ph = Gs, Ps
                torch.cuda.synchronize(dev)
                hG.barrier(channel=0, timeout_ms=20000)        # every rank's buffers are initialised
            except Exception as e:     # noqa: BLE001 - any failure of the peer mapping: keep the NCCL exchange
                if self.ex.rank == 0:
                    print(f"[lidar-nerf_b200] symmetric memory unavailable ({type(e).__name__}: {e}); NCCL exchange")
                self._peer = None
        if self.external_params:
            self.P = self.m = self.v = None
        elif self.ex.world > 1:
            # fp32 master weights and Adam moments exist only for this rank's shard (1/world of 3 x 54.8 MB)
            self.P = P[self.ex.lo:self.ex.hi].to(dev)
            self.G_shard = torch.zeros(self.ex.shard, dtype=torch.float32, device=dev)
            self.Ph_shard = torch.empty(self.ex.shard, dtype=torch.float16, device=dev)
        else:
            self.P = P.to(dev)
        if not self.external_params:
            self.m = torch.zeros_like(self.P)
            self.v = torch.zeros_like(self.P)
        self.table_h = self.Ph[:n_table].view(self.n_rows, c.level_dim)
        self.w_sigma_h = self.Ph[n_table:n_table + n_sigma]
        self.w_head_h = self.Ph[n_table + n_sigma:n]
        # bf16 MLPs (cfg.mlp_dtype): the kernels' `_bf16` builds read the weights from a separate bf16 copy of the MLP part
        # of the parameters (36 KB), refreshed inside the step from the fp32 master (one rank) or the fp16 shadow (data
        # parallel / external parameters); the table part of the shadow stays fp16
        self.bf16 = c.mlp_dtype == "bf16"
        if c.mlp_dtype not in ("fp16", "bf16"):
            raise ValueError(f"mlp_dtype={c.mlp_dtype!r}: choose fp16 or bf16")
        self.mlp_torch_dtype = torch.bfloat16 if self.bf16 else torch.float16
        self._sfx = "_bf16" if self.bf16 else ""
        if self.bf16:
            self.w_mlp_bf16 = torch.zeros(n_sigma + n_head, dtype=torch.bfloat16, device=dev)
            self.w_sigma_h = self.w_mlp_bf16[:n_sigma]
            self.w_head_h = self.w_mlp_bf16[n_sigma:]
            self._refresh_bf16_weights()
        if self.G is not None:
            self.set_grad_buffer(self.G)
        self.step_count = 0

        # ---- occupancy state -------------------------------------------------------
        H3 = c.grid_size ** 3
        self.density_grid = torch.zeros(c.cascade, H3, dtype=torch.float32, device=dev)
        self.prior_grid = torch.zeros(c.cascade, H3, dtype=torch.float32, device=dev)    # LiDAR free-space prior
        self.bitfield = torch.full((c.cascade * H3 // 8,), 255, dtype=torch.uint8, device=dev)
        self.counter = torch.zeros(4, dtype=torch.int32, device=dev)      # (samples, rays, live samples, -)
        self.mean_density = 0.0

        # ---- static per-ray buffers ------------------------------------------------------------------------
        N = self.N
        f = dict(dtype=torch.float32, device=dev)
        # one buffer for the step's inputs, [3, N, 3] = (rays_o | rays_d | gt): a batch prepared in this layout on the
        # host arrives with ONE copy (set_batch_packed); the three views are what the kernels read
        self.batch = torch.zeros(3, N, 3, **f)
        self.rays_o, self.rays_d, self.gt = self.batch[0], self.batch[1], self.batch[2]
        self.nears = torch.full((N,), c.min_near_lidar, **f)
        self.fars = self.nears * c.far_factor
        self.noises = torch.zeros(N, **f)
        self.t0 = torch.zeros(N, **f)
        self.rays = torch.zeros(N, 3, dtype=torch.int32, device=dev)
        self.ws = torch.zeros(N, **f)
        self.depth = torch.zeros(N, **f)
        self.image = torch.zeros(N, 2, **f)
        self.g_ws = torch.zeros(N, **f)
        self.g_depth = torch.zeros(N, **f)
        self.g_image = torch.zeros(N, 2, **f)
        self.loss_acc = torch.zeros(1, **f)
        two_sqrt3 = 2 * 1.7320508075688772
        self.dt_min = np.float32(two_sqrt3) / np.float32(c.max_steps)
        self.dt_max = np.float32(two_sqrt3) * np.float32(1 << (c.cascade - 1)) / np.float32(c.grid_size)

        # fused field kernels: per-ray direction terms instead of a per-sample [M, 96] head input
        self.fused = bool(c.fused_field) and self._fn("lnb_field_supported")(
            u32(self.enc_dim), u32(c.sigma_layers), u32(c.head_in_dim), u32(c.head_layers), u32(c.dir_code),
            u32(c.hidden_dim)) == 0
        if c.dir_encoding == "sh" and not self.fused:
            raise RuntimeError("dir_encoding='sh' is implemented by the fused field kernels only (fused_field=True, "
                               "SH degree 4, 64-wide 2-layer MLPs)")
        # gather + MLPs as one persistent kernel: the MLP weights travel as a pre-laid-out shared-memory image (TMA)
        self._fn("lnb_field_fused_weight_bytes").restype = C.c_size_t
        wbytes = int(self._fn("lnb_field_fused_weight_bytes")(u32(self.enc_dim), u32(c.sigma_layers), u32(c.head_in_dim),
                                                      u32(c.head_layers), u32(c.dir_code), u32(c.hidden_dim)))
        self.fused_gather = bool(c.fused_gather) and self.fused and wbytes > 0 and c.level_dim == 2
        if self.bf16 and not self.fused_gather:
            raise RuntimeError("mlp_dtype='bf16' needs the persistent forward kernel (fused_field and fused_gather, "
                               "level_dim 2): the stand-alone grid encoder writes fp16 features")
        self.wimage = torch.zeros(max(wbytes, 16), dtype=torch.uint8, device=dev) if self.fused_gather else None
        if self.fused_gather and c.l2_persist_table:
            # keep the fp16 table (27 MB at T = 2^19) resident in L2 across the step's streaming kernels
            lib.lnb_field_set_l2_window.argtypes = [C.c_void_p, C.c_size_t]
            rc = lib.lnb_field_set_l2_window(vp(self.table_h.data_ptr()), sz(self.n_table * 2))
            if rc != 0:                      # a cache hint only: without persisting L2 (MIG, MPS, ...) the kernel runs as before
                import warnings
                warnings.warn(f"lnb_field_set_l2_window: rc {rc}; the hash table is not pinned in L2")
        self.ray_enc = torch.zeros(N, c.head_in_dim, dtype=self.mlp_torch_dtype, device=dev)
        self.ray_bias = torch.zeros(N, c.hidden_dim, **f)

        # cross-step pipelining (graph mode, one rank): the captured step BEGINS with the Adam update of the previous
        # step's gradient on a side branch, next to the march of the new batch (the march does not read parameters)
        self._pipelined = bool(c.pipeline_adam) and self.ex.world == 1
        self._in_graph_body = False
        self._pending = False                      # G holds a complete gradient that Adam has not applied yet
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)    # lr, 1/bc1, 1/sqrt(bc2), grad scale, enable
        self.M = 0
        self._graph = None
        self._side = torch.cuda.Stream(device=dev)     # side branch of the step (per-ray direction terms)
        self._comm = torch.cuda.Stream(device=dev)     # data parallel: gradient exchange + sharded Adam
        self._graph_b = None
        self.graph_kernels = 0
        self._alloc_samples(sample_budget or N * 64)

    # ------------------------------------------------------------------------------------------------------
    def _fn(self, name):
        """C-ABI entry point of the MLP kernels for this engine's element type (`name` or `name_bf16`)."""
        return getattr(lib, name + self._sfx)

    def _refresh_bf16_weights(self):
        a, n = self.n_table, self.n_params
        src = self.P[a:n] if (self.P is not None and self.ex.world == 1) else self.Ph[a:n]
        self.w_mlp_bf16.copy_(src)

    def set_grad_buffer(self, G):
        """Flat fp32 gradient vector [hash table | density MLP | LiDAR head] the backward kernels accumulate into."""
        a, b, n = self.n_table, self.n_table + self.n_sigma, self.n_params
        self.G = G
        self.g_table = G[:a]
        self.g_sigma_w = G[a:b]
        self.g_head_w = G[b:n]

    def load_params(self, embeddings, w_sigma, w_head):
        """fp32 (or fp16) parameters owned by the caller -> the fp16 shadow the kernels read (three cast-copies)."""
        self.table_h.copy_(embeddings.detach().reshape(self.n_rows, self.cfg.level_dim))
        self.w_sigma_h.copy_(w_sigma.detach().reshape(-1))          # (fp16 shadow slice, or the bf16 copy in bf16 mode)
        self.w_head_h.copy_(w_head.detach().reshape(-1))

    # ------------------------------------------------------------------------------------------------------
    def _alloc_samples(self, M):
        M = max(128, (int(M) + 127) // 128 * 128)
        if M == self.M:
            return
        self.M = M
        self._graph = None
        dev, c = self.dev, self.cfg
        f = dict(dtype=torch.float32, device=dev)
        h = dict(dtype=self.mlp_torch_dtype, device=dev)          # MLP-typed rows (fp16, or bf16 with mlp_dtype="bf16")
        self.xyzs = torch.zeros(M, 3, **f)
        self.dirs = torch.zeros(M, 3, **f) if not self.fused else None
        self.ray_ids = torch.zeros(M, dtype=torch.int32, device=dev)
        self.live_idx = torch.zeros(M, dtype=torch.int32, device=dev)      # rows that can carry a gradient, compact
        self.deltas = torch.zeros(M, 2, **f)
        self.enc = torch.empty(M, self.enc_dim, **h)
        self.sig_out = torch.empty(M, 16, **h)
        self.fb_sigma = torch.empty(c.sigma_layers, M, c.hidden_dim, **h)
        self.sigma = torch.empty(M, **f)
        if not self.fused:
            self.head_in = torch.empty(M, c.head_in_dim, **h)
            self.head_out = torch.empty(M, 16, **h)
            self.g_head_out = torch.empty(M, 16, **h)
            self.g_head_in = torch.empty(M, c.head_in_dim, **h)
        self.fb_head = torch.empty(c.head_layers, M, c.hidden_dim, **h)
        self.rgb = torch.empty(M, 2, **f)
        self.g_sigma = torch.zeros(M, **f)
        self.g_rgb = torch.zeros(M, 2, **f)
        self.g_sig_out = torch.empty(M, 16, **h)
        self.g_enc = torch.empty(M, self.enc_dim, dtype=torch.float16, device=dev)   # always fp16: read by the scatter

    @staticmethod
    def _s():
        return vp(torch.cuda.current_stream().cuda_stream)

    # ------------------------------------------------------------------------------------------------------
    def _forward_backward(self):
        """Everything between 'rays are in the static buffers' and 'flat gradient is complete'."""
        self._fb_march()
        self._fb_field()

    def _fb_march(self):
        """The part of the step that does not read parameters: jitter noise and the occupancy march of the batch.  (The
        data-parallel step runs it next to the previous step's gradient exchange.)"""
        c, N, M, s = self.cfg, self.N, self.M, self._s()
        p = lambda t: vp(t.data_ptr())   # noqa: E731
        self.counter.zero_()
        if c.perturb:
            self.noises.uniform_(0, 1)
        else:
            self.noises.zero_()
        if not c.fused_composite:
            # march start per ray, for the absolute-depth term of the loss (raymarching.cu:375)
            torch.clamp(self.nears * c.dt_gamma, float(self.dt_min), float(self.dt_max), out=self.t0)
            torch.addcmul(self.nears, self.t0, self.noises, out=self.t0)
        _ck(lib.lnb_march_rays_train_ex(p(self.rays_o), p(self.rays_d), p(self.bitfield), f32(c.bound), f32(c.dt_gamma),
                                        u32(c.max_steps), u32(N), u32(c.cascade), u32(c.grid_size), u32(M),
                                        p(self.nears), p(self.fars), p(self.xyzs),
                                        p(self.dirs) if self.dirs is not None else vp(0), p(self.deltas), p(self.rays),
                                        p(self.counter), p(self.noises), p(self.ray_ids), s), "march_rays_train")

    def _fb_field(self):
        """Encoding, field network, compositing + loss, backward: everything that reads parameters."""
        self._fwd_field()
        self._composite_loss()
        self._bwd_field()

    def _fwd_field(self):
        """samples -> hash-grid features -> density MLP -> LiDAR head: sigma [M], rgb [M,2] (+ what backward needs)."""
        c, N, M, s = self.cfg, self.N, self.M, self._s()
        p = lambda t: vp(t.data_ptr())   # noqa: E731
        main = torch.cuda.current_stream()
        if self.fused:
            # per-ray direction terms depend only on rays_d and the head weights: a side branch next to the gather
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                if self.bf16:
                    self._refresh_bf16_weights()           # 18 k weights: fp32 master / fp16 shadow -> bf16 operand copy
                _ck(self._fn("lnb_field_ray_terms")(p(self.rays_d), p(self.w_head_h), u32(N), u32(c.dir_code),
                                            u32(c.head_in_dim), p(self.ray_enc), p(self.ray_bias), self._s()),
                    "ray_terms")
                if self.fused_gather:
                    _ck(self._fn("lnb_field_pack_weights")(p(self.w_sigma_h), p(self.w_head_h), u32(self.enc_dim),
                                                   u32(c.sigma_layers), u32(c.head_in_dim), u32(c.head_layers),
                                                   u32(c.dir_code), u32(c.hidden_dim), p(self.wimage), self._s()),
                        "pack_weights")
        # every per-sample kernel below reads the produced count from `counter` ON THE DEVICE and only touches
        # round_up(count, 128) rows, so M can be sized generously (no dropped rays) at no cost
        # (the extended march also zeroes the padding rows of the last tile)
        na = p(self.counter)
        if self._in_graph_body and self._pipelined:
            main.wait_stream(self._side)               # the pipelined Adam (same side stream) must be done before the gather
        compact = bool(c.compact_backward and c.fused_composite and self.fused)
        nl = vp(self.counter.data_ptr() + 8)          # counter[2]: live rows, counted by the compositing kernel
        if self.fused_gather:
            main.wait_stream(self._side)               # join: ray terms + weight image ready
            _ck(self._fn("lnb_field_fused_forward")(p(self.xyzs), p(self.table_h), p(self.offsets), u32(c.num_levels),
                                            u32(c.level_dim), f32(self.S), u32(c.base_resolution), f32(c.bound),
                                            p(self.wimage), p(self.ray_ids), p(self.ray_bias), u32(M),
                                            u32(c.sigma_layers), u32(c.head_in_dim), u32(c.head_layers),
                                            u32(c.dir_code), u32(c.hidden_dim), f32(c.density_scale), p(self.enc),
                                            p(self.fb_sigma), p(self.sig_out), p(self.sigma), p(self.fb_head),
                                            p(self.rgb), na, s), "field_fused_forward")
            return
        _ck(lib.lnb_grid_encode_forward_ex(p(self.xyzs), p(self.table_h), p(self.offsets), p(self.enc), u32(M), u32(3),
                                           u32(c.level_dim), u32(c.num_levels), f32(self.S), u32(c.base_resolution),
                                           vp(0), u32(0), i32(0), u32(0), i32(1), i32(1), f32(c.bound), na, s),
            "grid_fwd")
        if self.fused:
            main.wait_stream(self._side)               # join: ray terms ready
            _ck(self._fn("lnb_field_forward")(p(self.enc), p(self.w_sigma_h), p(self.w_head_h), p(self.ray_ids),
                                      p(self.ray_bias), u32(M), u32(self.enc_dim), u32(c.sigma_layers),
                                      u32(c.head_in_dim), u32(c.head_layers), u32(c.dir_code), u32(c.hidden_dim),
                                      f32(c.density_scale), p(self.fb_sigma), p(self.sig_out), p(self.sigma),
                                      p(self.fb_head), p(self.rgb), na, s), "field_forward")
        else:
            _ck(lib.lnb_ffmlp_forward_ex(p(self.enc), p(self.w_sigma_h), u32(M), u32(self.enc_dim), u32(16),
                                         u32(c.hidden_dim), u32(c.sigma_layers), u32(0), u32(6), p(self.fb_sigma),
                                         p(self.sig_out), na, s), "ffmlp_fwd(sigma)")
            _ck(lib.lnb_field_head_input(p(self.sig_out), p(self.dirs), u32(M), u32(c.freq_degree),
                                         u32(c.head_in_dim), f32(c.density_scale), p(self.sigma), p(self.head_in), na,
                                         s), "head_input")
            _ck(lib.lnb_ffmlp_forward_ex(p(self.head_in), p(self.w_head_h), u32(M), u32(c.head_in_dim), u32(16),
                                         u32(c.hidden_dim), u32(c.head_layers), u32(0), u32(6), p(self.fb_head),
                                         p(self.head_out), na, s), "ffmlp_fwd(head)")
            _ck(lib.lnb_field_head_rgb(p(self.head_out), u32(M), p(self.rgb), na, s), "head_rgb")

    def _compact(self):
        c = self.cfg
        return bool(c.compact_backward and c.fused_composite and self.fused)

    def composite_forward(self):
        """sigma, rgb -> per-ray weights_sum / depth (relative to the march start t0) / image (raymarching.cu:578-676)."""
        c = self.cfg
        rm.composite_rays_train_forward_ex(self.sigma, self.rgb, self.deltas, self.rays, self.M, self.N, c.T_thresh, 2,
                                           self.ws, self.depth, self.image)

    def composite_backward(self):
        """per-ray gradients in g_ws / g_depth / g_image -> g_sigma, g_rgb (with the depth gradient the reference drops,
        raymarching.py:329-330)."""
        c = self.cfg
        self.g_sigma.zero_()
        self.g_rgb.zero_()
        rm.composite_rays_train_backward_ex(self.g_ws, self.g_depth, self.g_image, self.sigma, self.rgb,
                                            self.deltas, self.rays, self.ws, self.depth, self.image, self.M, self.N,
                                            c.T_thresh, 2, self.g_sigma, self.g_rgb)

    def _composite_loss(self):
        """compositing forward + LiDAR loss + compositing backward: one kernel (fused_composite) or the per-op chain."""
        c, N, M, s = self.cfg, self.N, self.M, self._s()
        p = lambda t: vp(t.data_ptr())   # noqa: E731
        na = p(self.counter)
        compact = self._compact()
        nl = vp(self.counter.data_ptr() + 8)          # counter[2]: live rows, counted by the compositing kernel
        if c.fused_composite and c.patch_loss:
            # a loss that couples neighbouring rays: forward of every ray -> loss + per-ray gradients -> backward
            _ck(lib.lnb_lidar_composite_forward(p(self.sigma), p(self.rgb), p(self.deltas), p(self.rays), p(self.gt),
                                                p(self.nears), p(self.noises), f32(c.dt_gamma), u32(c.max_steps),
                                                u32(c.cascade), u32(c.grid_size), u32(M), u32(N), f32(c.T_thresh),
                                                p(self.ws), p(self.depth), p(self.image), p(self.t0), s),
                "lidar_composite_forward")
            _ck(lib.lnb_lidar_loss_ex(p(self.ws), p(self.depth), p(self.image), p(self.gt), p(self.t0), u32(N),
                                      f32(c.alpha_d), f32(c.alpha_r), f32(c.alpha_i), f32(c.loss_scale),
                                      u32(c.patch_size[0]), u32(c.patch_size[1]), f32(c.alpha_grad),
                                      f32(1.0 / c.min_near_lidar), f32(c.grad_clip), p(self.g_ws), p(self.g_depth),
                                      p(self.g_image), p(self.loss_acc), s), "lidar_loss_ex")
            _ck(lib.lnb_lidar_composite_backward(p(self.g_ws), p(self.g_depth), p(self.g_image), p(self.sigma), p(self.rgb),
                                                 p(self.deltas), p(self.rays), p(self.gt), p(self.nears), p(self.noises),
                                                 f32(c.dt_gamma), u32(c.max_steps), u32(c.cascade), u32(c.grid_size), na,
                                                 u32(M), u32(N), f32(c.T_thresh), p(self.ws), p(self.depth), p(self.image),
                                                 p(self.g_sigma), p(self.g_rgb), p(self.live_idx) if compact else vp(0),
                                                 nl if compact else vp(0), s), "lidar_composite_backward")
        elif c.fused_composite:
            _ck(lib.lnb_lidar_composite_step(p(self.sigma), p(self.rgb), p(self.deltas), p(self.rays), p(self.gt),
                                             p(self.nears), p(self.noises), f32(c.dt_gamma), u32(c.max_steps),
                                             u32(c.cascade), u32(c.grid_size), na, u32(M), u32(N), f32(c.T_thresh),
                                             f32(c.alpha_d), f32(c.alpha_r), f32(c.alpha_i), f32(c.loss_scale),
                                             p(self.ws), p(self.depth), p(self.image), p(self.t0), p(self.g_sigma),
                                             p(self.g_rgb), p(self.loss_acc), p(self.live_idx) if compact else vp(0),
                                             nl if compact else vp(0), s), "lidar_composite_step")
        else:
            self.composite_forward()
            _ck(lib.lnb_lidar_loss_ex(p(self.ws), p(self.depth), p(self.image), p(self.gt), p(self.t0), u32(N),
                                      f32(c.alpha_d), f32(c.alpha_r), f32(c.alpha_i), f32(c.loss_scale),
                                      u32(c.patch_size[0]), u32(c.patch_size[1]), f32(c.alpha_grad if c.patch_loss else 0.0),
                                      f32(1.0 / c.min_near_lidar), f32(c.grad_clip), p(self.g_ws), p(self.g_depth),
                                      p(self.g_image), p(self.loss_acc), s), "lidar_loss")
            self.composite_backward()

    def _bwd_field(self):
        """g_sigma, g_rgb -> gradient of the LiDAR head, the density MLP and the hash table (accumulated into G)."""
        c, N, M, s = self.cfg, self.N, self.M, self._s()
        p = lambda t: vp(t.data_ptr())   # noqa: E731
        na = p(self.counter)
        compact = self._compact()
        nl = vp(self.counter.data_ptr() + 8)
        if c.late_grad_zero:
            # The 55 MB fp32 gradient table is cleared HERE, a few microseconds before the scatter-add, instead of by
            # Adam half a millisecond (and ~1 GB of other traffic) earlier: the zeroed lines are still in the 126 MB L2
            # when the atomics arrive, so they do not have to be fetched back from HBM one 32-byte sector at a time.
            self.G[self.n_table:].zero_()
        if compact:
            # backward on the live rows only: g_sig_out / g_enc are in compact order, everything saved by the forward
            # pass is read through live_idx
            li = p(self.live_idx)
            _ck(self._fn("lnb_field_head_backward_rows")(p(self.g_rgb), p(self.rgb), p(self.g_sigma), p(self.sig_out),
                                                 p(self.ray_ids), p(self.ray_enc), p(self.w_head_h), p(self.fb_head),
                                                 u32(M), u32(c.head_in_dim), u32(c.head_layers), u32(c.dir_code),
                                                 u32(c.hidden_dim), f32(c.density_scale), p(self.g_sig_out),
                                                 p(self.g_head_w), li, nl, s), "field_head_backward_rows")
            _ck(self._fn("lnb_ffmlp_backward_accumulate_rows")(p(self.g_sig_out), p(self.enc), p(self.w_sigma_h),
                                                       p(self.fb_sigma), u32(M), u32(self.enc_dim), u32(16),
                                                       u32(c.hidden_dim), u32(c.sigma_layers), u32(0), u32(6), i32(1),
                                                       p(self.g_enc), p(self.g_sigma_w), li, nl, s),
                "ffmlp_bwd_rows(sigma)")
            if c.late_grad_zero:
                # (Measured and rejected: clearing the table on a forked branch UNDER the density-MLP backward hides the
                # 10 us memset but costs 20 us - the zeroed lines must be the last thing that entered L2 before the scatter.)
                self.g_table.zero_()
            _ck(lib.lnb_grid_encode_backward_rows(p(self.g_enc), p(self.xyzs), p(self.table_h), p(self.offsets),
                                                  p(self.g_table), u32(M), u32(3), u32(c.level_dim), u32(c.num_levels),
                                                  f32(self.S), u32(c.base_resolution), u32(0), i32(0), u32(0), i32(1),
                                                  f32(c.bound), i32(1), li, nl, s), "grid_bwd_rows")
            return
        if self.fused:
            _ck(self._fn("lnb_field_head_backward")(p(self.g_rgb), p(self.rgb), p(self.g_sigma), p(self.sig_out),
                                            p(self.ray_ids), p(self.ray_enc), p(self.w_head_h), p(self.fb_head),
                                            u32(M), u32(c.head_in_dim), u32(c.head_layers), u32(c.dir_code),
                                            u32(c.hidden_dim), f32(c.density_scale), p(self.g_sig_out),
                                            p(self.g_head_w), na, s), "field_head_backward")
        else:
            _ck(lib.lnb_field_head_out_grad(p(self.g_rgb), p(self.rgb), u32(M), p(self.g_head_out), na, s),
                "head_out_grad")
            _ck(lib.lnb_ffmlp_backward_accumulate(p(self.g_head_out), p(self.head_in), p(self.w_head_h),
                                                  p(self.fb_head), u32(M), u32(c.head_in_dim), u32(16),
                                                  u32(c.hidden_dim), u32(c.head_layers), u32(0), u32(6), i32(1),
                                                  p(self.g_head_in), p(self.g_head_w), na, s), "ffmlp_bwd(head)")
            _ck(lib.lnb_field_sigma_out_grad(p(self.g_sigma), p(self.sig_out), p(self.g_head_in), u32(M),
                                             u32(c.head_in_dim), u32(c.freq_degree), f32(c.density_scale),
                                             p(self.g_sig_out), na, s), "sigma_out_grad")
        _ck(self._fn("lnb_ffmlp_backward_accumulate")(p(self.g_sig_out), p(self.enc), p(self.w_sigma_h), p(self.fb_sigma),
                                              u32(M), u32(self.enc_dim), u32(16), u32(c.hidden_dim),
                                              u32(c.sigma_layers), u32(0), u32(6), i32(1), p(self.g_enc),
                                              p(self.g_sigma_w), na, s), "ffmlp_bwd(sigma)")
        if c.late_grad_zero:
            self.g_table.zero_()
        _ck(lib.lnb_grid_encode_backward_ex(p(self.g_enc), p(self.xyzs), p(self.table_h), p(self.offsets),
                                            p(self.g_table), u32(M), u32(3), u32(c.level_dim), u32(c.num_levels),
                                            f32(self.S), u32(c.base_resolution), vp(0), vp(0), u32(0), i32(0), u32(0),
                                            i32(1), i32(1), f32(c.bound), i32(1), na, s), "grid_bwd")

    def _optimizer(self, lr=None):
        """Adam on the gradient currently in G, now.  One rank: a single fused pass over the whole flat vector.  Data
        parallel: reduce-scatter the fp32 gradient, update only this rank's shard (fp32 master + moments live only
        here), all-gather the fp16 shadow."""
        c = self.cfg
        self.step_count += 1
        self._pending = False
        lr = c.lr if lr is None else lr
        if self.ex.world == 1:
            adam_step(self.P, self.G, self.m, self.v, self.Ph, lr, c.beta1, c.beta2, c.eps, self.step_count,
                      grad_scale=1.0 / c.loss_scale, zero_grad=not c.late_grad_zero)
            return
        if self._peer is not None:
            # all ranks' gradients complete -> [peer reduce-scatter + Adam + peer all-gather] -> all shadows complete
            pr = self._peer
            pr["hG"].barrier(channel=0, timeout_ms=20000)
            if pr["mc"] is not None:
                _ck(lib.lnb_dp_adam_exchange_mc(vp(pr["mc"][0]), vp(pr["mc"][1]), vp(self.P.data_ptr()),
                                                vp(self.m.data_ptr()), vp(self.v.data_ptr()), sz(self.ex.lo),
                                                sz(self.ex.shard), f32(lr), f32(c.beta1), f32(c.beta2), f32(c.eps),
                                                f32(1.0 - c.beta1 ** self.step_count),
                                                f32(1.0 - c.beta2 ** self.step_count),
                                                f32(dp.grad_scale(c.loss_scale)), self._s()), "dp_adam_exchange_mc")
            else:
                _ck(lib.lnb_dp_adam_exchange(pr["g"], pr["h"], u32(self.ex.world), vp(self.P.data_ptr()),
                                             vp(self.m.data_ptr()), vp(self.v.data_ptr()), sz(self.ex.lo),
                                             sz(self.ex.shard), f32(lr), f32(c.beta1), f32(c.beta2), f32(c.eps),
                                             f32(1.0 - c.beta1 ** self.step_count),
                                             f32(1.0 - c.beta2 ** self.step_count),
                                             f32(dp.grad_scale(c.loss_scale)), self._s()), "dp_adam_exchange")
            pr["hG"].barrier(channel=1, timeout_ms=20000)
            if not c.late_grad_zero:
                self.G.zero_()
            return
        self.ex.reduce_scatter(self.G, self.G_shard)
        if not c.late_grad_zero:
            self.G.zero_()
        adam_step(self.P, self.G_shard, self.m, self.v, self.Ph_shard, lr, c.beta1, c.beta2, c.eps, self.step_count,
                  grad_scale=dp.grad_scale(c.loss_scale), zero_grad=False)
        self.ex.all_gather(self.Ph, self.Ph_shard)

    def flush(self):
        """Settle what the asynchronous step schedules still owe: the update of the pipelined graph step (the gradient of
        the last step) and, data parallel, the gradient exchange running on the communication stream.  Call before
        reading the parameters, refreshing the density grid, or mixing in eager steps; cheap when nothing is pending."""
        if self._pending:
            self._optimizer()
        if self.ex.world > 1:
            torch.cuda.current_stream().wait_stream(self._comm)

    def _set_hyper(self, enable, lr=None):
        c = self.cfg
        t = max(self.step_count, 1)
        _ck(lib.lnb_adam_set_hyper(vp(self.hyper.data_ptr()), f32(c.lr if lr is None else lr),
                                   f32(1.0 - c.beta1 ** t), f32(1.0 - c.beta2 ** t), f32(1.0 / c.loss_scale),
                                   i32(1 if enable else 0), self._s()), "adam_set_hyper")

    def _graph_body(self):
        """What the CUDA graph holds: [Adam of the previous gradient || march] -> forward -> backward."""
        c = self.cfg
        self._in_graph_body = True
        try:
            if self._pipelined:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    # hyper-parameters come from device memory (written by _set_hyper before every replay); the per-ray
                    # direction terms of _forward_backward follow on the same side stream, after the update
                    _ck(lib.lnb_adam_step_dev(vp(self.P.data_ptr()), vp(self.G.data_ptr()), vp(self.m.data_ptr()),
                                              vp(self.v.data_ptr()), vp(self.Ph.data_ptr()), sz(self.P.numel()),
                                              f32(c.beta1), f32(c.beta2), f32(c.eps), vp(self.hyper.data_ptr()),
                                              i32(0 if c.late_grad_zero else 1), self._s()), "adam_step_dev")
            self._forward_backward()
        finally:
            self._in_graph_body = False

    def set_batch(self, rays_o, rays_d, gt):
        """Device tensors [N,3] each (gt = ray-drop, intensity, depth) -> static buffers."""
        self.rays_o.copy_(rays_o.reshape(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d.reshape(-1, 3), non_blocking=True)
        self.gt.copy_(gt.reshape(-1, 3), non_blocking=True)

    def set_batch_packed(self, batch):
        """[3, N, 3] tensor (rays_o | rays_d | gt), device or pinned host memory -> static buffers, one async copy."""
        self.batch.copy_(batch, non_blocking=True)

    # ---- host batches: double-buffered staging so that the H2D copy of step i+1 runs under the compute of step i --------
    def stage_batch(self, host_batch):
        """Start the asynchronous copy of a pinned host batch [3, N, 3] (rays_o | rays_d | gt) into the next staging slot
        on the engine's copy stream.  Pair with `set_batch_staged()`; call it for batch i+1 right after launching step i."""
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(self.batch) for _ in range(2)]
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [None, None]
            self._copy_stream = torch.cuda.Stream(device=self.dev)
            self._stage_w = 0
            self._stage_r = 0
        k = self._stage_w & 1
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[k] is not None:
                self._copy_stream.wait_event(self._stage_free[k])      # the step that read this slot has consumed it
            self._stage[k].copy_(host_batch, non_blocking=True)
            self._stage_ready[k].record(self._copy_stream)
        self._stage_w += 1

    def set_batch_staged(self):
        """Main stream: wait for the oldest staged batch and move it (device to device, 147 KB) into the static buffers."""
        k = self._stage_r & 1
        main = torch.cuda.current_stream()
        main.wait_event(self._stage_ready[k])
        self.batch.copy_(self._stage[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        self._stage_free[k] = ev
        self._stage_r += 1

    @torch.no_grad()
    def render(self, rays_o, rays_d, perturb=False):
        """Forward only, for evaluation: occupancy march (no jitter by default) -> field -> compositing of `rays_o/rays_d`
        [n,3] (any n; processed in chunks of the engine's ray count) with the CURRENT parameters and occupancy grid.
        Returns (weights_sum [n], absolute depth [n], image [n,2]) - what `NeRFRenderer.render` returns for the LiDAR
        branch (renderer.py:268-297)."""
        self.flush()
        c = self.cfg
        n = rays_o.shape[0]
        ws, depth, image = (torch.empty(n, device=self.dev), torch.empty(n, device=self.dev),
                            torch.empty(n, 2, device=self.dev))
        keep = (c.perturb, self.rays_o.clone(), self.rays_d.clone())
        c.perturb = bool(perturb)
        try:
            for lo in range(0, n, self.N):
                hi = min(lo + self.N, n)
                k = hi - lo
                self.rays_o[:k].copy_(rays_o[lo:hi])
                self.rays_d[:k].copy_(rays_d[lo:hi])
                if k < self.N:                       # pad the last chunk with copies of its first ray
                    self.rays_o[k:].copy_(rays_o[lo:lo + 1].expand(self.N - k, 3))
                    self.rays_d[k:].copy_(rays_d[lo:lo + 1].expand(self.N - k, 3))
                self._fb_march()
                if c.fused_composite:                # (t0 is otherwise produced by the fused compositing kernel)
                    torch.clamp(self.nears * c.dt_gamma, float(self.dt_min), float(self.dt_max), out=self.t0)
                    torch.addcmul(self.nears, self.t0, self.noises, out=self.t0)
                self._fwd_field()
                self.composite_forward()
                ws[lo:hi] = self.ws[:k]
                depth[lo:hi] = torch.addcmul(self.depth, self.t0, self.ws)[:k]
                image[lo:hi] = self.image[:k]
        finally:
            c.perturb = keep[0]
            self.rays_o.copy_(keep[1])
            self.rays_d.copy_(keep[2])
        return ws, depth, image

    def set_batch_from_pixels(self, pose, inds, image, H, W, fov_up, fov):
        """The reference's per-step collate on the device (kitti360_dataset.py:123-159): `pose` [4,4] fp32 (lidar2world,
        already offset/scaled), `inds` [N] int32 flat pixel indices, `image` [H*W, 3] fp32 (ray-drop, intensity,
        depth * scale) of a frame resident on the device -> rays_o / rays_d / gt in the static buffers, one kernel."""
        _ck(lib.lnb_lidar_batch(vp(pose.data_ptr()), vp(inds.data_ptr()), vp(image.data_ptr()), u32(self.N), u32(H), u32(W),
                                f32(fov_up), f32(fov), vp(self.rays_o.data_ptr()), vp(self.rays_d.data_ptr()),
                                vp(self.gt.data_ptr()), self._s()), "lidar_batch")

    def train_step(self, use_graph=True):
        """One optimiser step on the batch currently in the static buffers.  In graph mode on one rank the update is
        applied at the START of the next step (next to its march) - call flush() before reading parameters."""
        interval = self.cfg.grid_update_interval
        if use_graph and self._pipelined:
            if self._graph is None:
                self.flush()
                self._capture()
            owed = self._pending
            if owed:
                self.step_count += 1
            self._set_hyper(owed)
            self._graph.replay()          # [Adam(previous gradient) || march] ... grid backward: one graph launch
            self._pending = True
            if interval > 0 and (self.step_count + 1) % interval == 0:
                self.flush()              # the refresh evaluates the network: it needs this step's update
                self.update_density_grid()
            return
        if use_graph and self.ex.world > 1 and self.cfg.overlap_exchange:
            # data parallel: [march of THIS batch] runs while the previous step's gradient exchange is still in flight on
            # the communication stream; the rest of the step waits for the all-gathered parameters
            if self._graph is None:
                self._capture()
            main = torch.cuda.current_stream()
            self._graph.replay()                        # graph A: noise + march (reads no parameters)
            main.wait_stream(self._comm)
            self._graph_b.replay()                      # graph B: gather ... scatter
            self._comm.wait_stream(main)
            with torch.cuda.stream(self._comm):
                self._optimizer()                       # reduce-scatter -> Adam on the shard -> all-gather
            if interval > 0 and self.step_count % interval == 0:
                main.wait_stream(self._comm)
                self.update_density_grid()
            return
        self.flush()
        if use_graph:
            if self._graph is None:
                self._capture()
            self._graph.replay()          # march ... grid backward: one graph launch
        else:
            self._forward_backward()
        self._optimizer()                 # Adam (+ the data-parallel exchange)
        if interval > 0 and self.step_count % interval == 0:
            self.update_density_grid()

    def _capture(self):
        # warm up on a side stream (module loads, cudaFuncSetAttribute) before capturing
        if self._pipelined:
            self._set_hyper(False)        # the captured Adam is a no-op during the warm-up run
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            g_backup = self.G.clone()
            self._graph_body()
            self.G.copy_(g_backup)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        # thread_local: other threads (e.g. NCCL's watchdog) may legally touch the CUDA API during the capture
        n0 = launch_count()
        if self.ex.world > 1 and self.cfg.overlap_exchange:
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, capture_error_mode="thread_local"):
                self._fb_march()
            with torch.cuda.graph(gb, capture_error_mode="thread_local"):
                self._fb_field()
            self._graph, self._graph_b = ga, gb
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._graph_body()
            self._graph = g
        # kernels of liblnb200.so that ONE replay of the captured step executes (bench.py's `gpu_launches`)
        self.graph_kernels = launch_count() - n0

    # ------------------------------------------------------------------------------------------------------
    def samples_last_step(self):
        """(samples produced, rays marched) of the last step - one D2H sync; not called inside the timed loop."""
        cnt = self.counter.cpu()
        return int(cnt[0]), int(cnt[1])

    def fit_sample_budget(self, headroom=1.5):
        """Size M from the count of the last step (the role of the reference's mean_count, raymarching.py:228-233).
        Kernels only process the produced rows, so generous headroom costs memory, not time."""
        produced, _ = self.samples_last_step()
        want = max(128, int(produced * headroom))
        if produced > 0.9 * self.M or want < 0.4 * self.M:
            self._alloc_samples(want)
        return self.M

    LOSS_RING = 4

    def read_loss_async(self):
        """Device->host read of the step's loss without draining the launch queue: the 4 bytes are copied into pinned
        memory behind the step (stream-ordered) and the accumulator is cleared; returns the loss copied LOSS_RING - 1
        calls ago (None until then), which has long arrived - the host never waits for the step it has just launched and
        stays a few steps ahead of the device.  `read_loss_last()` collects the newest one."""
        R = self.LOSS_RING
        if not hasattr(self, "_loss_pin"):
            self._loss_pin = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(R)]
            self._loss_ev = [None] * R
            self._loss_slot = 0
        k = self._loss_slot
        old_ev = self._loss_ev[k]
        old = None
        if old_ev is not None:          # the slot about to be overwritten holds the loss of R calls ago: hand it out
            old_ev.synchronize()
            old = float(self._loss_pin[k][0])
        self._loss_pin[k].copy_(self.loss_acc, non_blocking=True)
        self.loss_acc.zero_()
        ev = torch.cuda.Event()
        ev.record()
        self._loss_ev[k] = ev
        self._loss_slot = (k + 1) % R
        return old

    def read_loss_last(self):
        if not hasattr(self, "_loss_pin"):
            return None
        k = (self._loss_slot - 1) % self.LOSS_RING
        if self._loss_ev[k] is None:
            return None
        self._loss_ev[k].synchronize()
        return float(self._loss_pin[k][0])

    def read_loss(self, reset=True):
        v = float(self.loss_acc.item())
        if reset:
            self.loss_acc.zero_()
        return v

    # ---- occupancy grid --------------------------------------------------------------------------------------
    def cell_centers(self, cas, idx=None, jitter=True):
        """World-space centres (optionally jittered inside the cell) of the cells `idx` (Morton indices; all H^3 cells
        if None) of one cascade - Morton order is the layout packbits / the march expect."""
        from ..raymarching import morton3D_invert
        c = self.cfg
        H = c.grid_size
        if idx is None:
            idx = torch.arange(H ** 3, dtype=torch.int32, device=self.dev)
        coords = morton3D_invert(idx.int()).float()                  # [n, 3] integer cell coordinates
        xyz = 2 * coords / (H - 1) - 1                               # [-1, 1] (upstream convention)
        bound = min(2.0 ** cas, c.bound)
        half = bound / H
        xyz = xyz * (bound - half)
        if jitter:
            xyz = xyz + (torch.rand_like(xyz) * 2 - 1) * half
        return xyz

    @torch.no_grad()
    def query_density(self, xyz):
        """sigma at arbitrary points [B,3] (inference kernels; B padded to 128)."""
        c = self.cfg
        B = xyz.shape[0]
        Bp = (B + 127) // 128 * 128
        pts = torch.zeros(Bp, 3, dtype=torch.float32, device=self.dev)
        pts[:B] = xyz
        enc = torch.empty(Bp, self.enc_dim, dtype=torch.float16, device=self.dev)
        s = self._s()
        p = lambda t: vp(t.data_ptr())   # noqa: E731
        _ck(lib.lnb_grid_encode_forward_ex(p(pts), p(self.table_h), p(self.offsets), p(enc), u32(Bp), u32(3),
                                           u32(c.level_dim), u32(c.num_levels), f32(self.S), u32(c.base_resolution),
                                           vp(0), u32(0), i32(0), u32(0), i32(1), i32(1), f32(c.bound), vp(0), s),
            "grid_fwd")
        out = torch.empty(Bp, 16, dtype=self.mlp_torch_dtype, device=self.dev)
        if self.bf16:
            self._refresh_bf16_weights()
            enc = enc.to(torch.bfloat16)
            _ck(lib.lnb_ffmlp_inference_bf16(p(enc), p(self.w_sigma_h), u32(Bp), u32(self.enc_dim), u32(16),
                                             u32(c.hidden_dim), u32(c.sigma_layers), u32(0), u32(6), vp(0), p(out), s),
                "ffmlp_inference_bf16")
        else:
            ff.ffmlp_inference(enc, self.w_sigma_h, Bp, self.enc_dim, 16, c.hidden_dim, c.sigma_layers, 0, 6, None, out)
        return torch.exp(out[:B, 0].float()) * c.density_scale

    @torch.no_grad()
    def update_density_grid(self, decay=0.95, full=None):
        """EMA-max refresh of the density grid from the current network + packbits, merged
        with the LiDAR prior grid (cells a GT return falls into stay occupied).  The first 16 refreshes visit every
        cell; the steady-state partial refresh (H^3/4 uniform + H^3/4 occupied cells per cascade) runs WITHOUT any host
        synchronisation and is replayed as one CUDA graph."""
        from ..raymarching import packbits
        c = self.cfg
        if self.ex.world > 1:
            torch.cuda.current_stream().wait_stream(self._comm)   # the network is evaluated: parameters must be settled
        n_updates = self.step_count // max(c.grid_update_interval, 1)
        full = (n_updates <= 16) if full is None else full
        if not full:
            return self._refresh_partial(decay)
        for cas in range(c.cascade):
            sig = self.query_density(self.cell_centers(cas))
            self.density_grid[cas] = torch.maximum(self.density_grid[cas] * decay, sig)
        merged = torch.maximum(self.density_grid, self.prior_grid)
        self.mean_density = float(merged.clamp(min=0).mean().item())
        thresh = min(self.mean_density, c.density_thresh)
        packbits(merged, thresh, self.bitfield)

    def _refresh_partial_body(self, decay):
        """Partial refresh, device-only: per cascade H^3/4 uniformly random cells + H^3/4 cells drawn uniformly from the
        currently occupied ones (the k-th occupied cell via a prefix sum of the occupancy mask + binary search instead of
        torch.nonzero, whose result size would have to travel to the host), re-evaluated with the current network."""
        c = self.cfg
        H3 = c.grid_size ** 3
        nq = H3 // 4
        dev = self.dev
        for cas in range(c.cascade):
            cur = self.density_grid[cas]
            mask = torch.maximum(cur, self.prior_grid[cas]) > 0
            cs = torch.cumsum(mask, 0, dtype=torch.int32)
            total = cs[-1:]
            rnd = torch.randint(0, H3, (nq,), device=dev)
            k = (torch.rand(nq, device=dev) * total).to(torch.int32)
            k = torch.minimum(k, torch.clamp(total - 1, min=0))
            occ = torch.searchsorted(cs, k + 1)
            occ = torch.where(total > 0, occ, rnd).clamp_(max=H3 - 1)
            sel = torch.cat([rnd, occ])
            sig = self.query_density(self.cell_centers(cas, sel))
            # duplicates in `sel` resolve to one of the candidate values, like torch-ngp's indexed assignment
            cur[sel] = torch.maximum(cur[sel] * decay, sig)
        merged = torch.maximum(self.density_grid, self.prior_grid)
        self._mean_density_dev.copy_(merged.clamp_(min=0).mean().reshape(1))
        _ck(lib.lnb_packbits_dev(vp(merged.data_ptr()), u32(merged.numel() // 8), vp(self._mean_density_dev.data_ptr()),
                                 f32(c.density_thresh), vp(self.bitfield.data_ptr()), self._s()), "packbits_dev")

    def _refresh_partial(self, decay):
        if not hasattr(self, "_mean_density_dev"):
            self._mean_density_dev = torch.zeros(1, dtype=torch.float32, device=self.dev)
            self._refresh_graph = None
            self._refresh_decay = None
        if os.environ.get("LNB_REFRESH_GRAPH", "1") == "0":
            return self._refresh_partial_body(decay)
        if self._refresh_graph is None or self._refresh_decay != decay:
            # warm up on a side stream (allocator, lazy loads), then capture; the refresh draws its random numbers
            # through torch's graph-safe generator state
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._refresh_partial_body(decay)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._refresh_partial_body(decay)
            self._refresh_graph, self._refresh_decay = g, decay
            return                               # (the warm-up run above was this call's refresh)
        self._refresh_graph.replay()

    @torch.no_grad()
    def seed_occupancy_from_points(self, points, dilate=1, value=1e4):
        """LiDAR prior: mark the cells containing GT returns (and their `dilate`-neighbourhood) as occupied in every
        cascade that contains them.  points: [P,3] world coordinates (already scaled into [-bound, bound])."""
        from ..raymarching import morton3D, packbits
        c = self.cfg
        H = c.grid_size
        self.prior_grid.zero_()
        offs = torch.stack(torch.meshgrid(*([torch.arange(-dilate, dilate + 1, device=self.dev)] * 3), indexing="ij"),
                           -1).reshape(-1, 3)
        for cas in range(c.cascade):
            bound = min(2.0 ** cas, c.bound)
            inside = (points.abs() <= bound).all(-1)
            pts = points[inside]
            if pts.numel() == 0:
                continue
            cell = torch.clamp((0.5 * (pts / bound + 1) * H).long(), 0, H - 1)
            cell = (cell[:, None, :] + offs[None]).reshape(-1, 3).clamp(0, H - 1)
            cell = torch.unique(cell, dim=0)
            idx = morton3D(cell.int()).long()
            self.prior_grid[cas, idx] = value
        merged = torch.maximum(self.density_grid, self.prior_grid)
        self.mean_density = float(merged.clamp(min=0).mean().item())
        packbits(merged, min(self.mean_density, c.density_thresh), self.bitfield)
