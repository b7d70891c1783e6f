"""TEST/BENCH INFRASTRUCTURE - not product code.

One training step of the LiDAR field assembled from the UNMODIFIED reference CUDA extensions compiled into
oracle/_ref/ (oracle/build_ref.py): `_raymarching`, `_gridencoder`, `_freqencoder`, `_ffmlp` called through their
pybind signatures the way the reference's Python wrappers call them:

  * raymarching.py:171-289 - fresh zero-filled xyzs/dirs/deltas and counter for every march call;
  * grid.py:24-138        - fp16 copy of the fp32 embedding table per forward, [L,B,C] output permuted to [B,L*C],
                            `zeros_like(embeddings)` gradient per backward;
  * ffmlp.py:14-164       - fp16 copy of the fp32 weights per forward, fresh forward/backward buffers and zero
                            gradient buffers per call, split-K workspace allocated once;
  * activation.py:6-20, network.py:162-237 - trunc_exp / sigmoid / concat done with torch ops;
  * torch.optim.Adam over the fp32 parameters (main_lidarnerf.py:389-391).

It is the GPU-side counterpart of oracle/field_step.py: same wiring as lidar_nerf_b200.nerf.engine, reference kernels
instead of ours.  Used for (1) an end-to-end loss parity check against the engine on identical parameters and rays
and (2) `bench.py --impl reference-cuda`, the "reference raymarching/ffmlp build" timing of SURVEY.md section 8(d).

Differences that are the reference's own: its composite backward has no depth-gradient input, so
the depth loss reaches sigma only through weights_sum; the march needs max_steps * N worst-case rows unless a
mean_count is supplied - we give it the engine's sample budget M (generous to the reference: no 134 MB memset).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import refcuda  # noqa: E402


class RefCudaStep:
    def __init__(self, eng):
        """Takes geometry, parameters and occupancy bitfield from a lidar_nerf_b200 LidarFieldEngine."""
        self.rm, self.ge = refcuda.load("_raymarching"), refcuda.load("_gridencoder")
        self.fe, self.ff = refcuda.load("_freqencoder"), refcuda.load("_ffmlp")
        missing = [n for n, m in (("_raymarching", self.rm), ("_gridencoder", self.ge), ("_freqencoder", self.fe),
                                  ("_ffmlp", self.ff)) if m is None]
        if missing:
            raise RuntimeError(f"reference CUDA extensions not built into oracle/_ref: {missing}")
        c = eng.cfg
        self.cfg, self.N, self.M, self.dev = c, eng.N, eng.M, eng.dev
        self.offsets, self.S, self.n_rows, self.enc_dim = eng.offsets, eng.S, eng.n_rows, eng.enc_dim
        full = eng.Ph[:eng.n_params].float() if eng.ex.world > 1 else eng.P[:eng.n_params].clone()
        a, b = eng.n_table, eng.n_table + eng.n_sigma
        self.embeddings = full[:a].view(self.n_rows, c.level_dim).clone().requires_grad_(False)
        self.w_sigma = full[a:b].clone()
        self.w_head = full[b:].clone()
        for p in (self.embeddings, self.w_sigma, self.w_head):
            p.grad = torch.zeros_like(p)
        self.opt = torch.optim.Adam([self.embeddings, self.w_sigma, self.w_head], lr=c.lr, betas=(c.beta1, c.beta2),
                                    eps=c.eps)
        self.bitfield = eng.bitfield
        self.ff.allocate_splitk(max(c.sigma_layers, c.head_layers) + 1)
        self.nfreq = 3 + 6 * c.freq_degree

    # one optimiser step; rays_o/rays_d [N,3], gt [N,3] (ray-drop mask, intensity, depth), noises [N]
    def step(self, rays_o, rays_d, gt, noises, apply_adam=True):
        c, N, M, dev = self.cfg, self.N, self.M, self.dev
        rm, ge, fe, ff = self.rm, self.ge, self.fe, self.ff
        f32, f16 = torch.float32, torch.float16
        nears = torch.full((N,), c.min_near_lidar, device=dev, dtype=f32)
        fars = nears * c.far_factor
        # ---- march (fresh zeroed outputs per call, as the reference wrapper does) ----
        xyzs = torch.zeros(M, 3, device=dev, dtype=f32)
        dirs = torch.zeros(M, 3, device=dev, dtype=f32)
        deltas = torch.zeros(M, 2, device=dev, dtype=f32)
        rays = torch.empty(N, 3, device=dev, dtype=torch.int32)
        counter = torch.zeros(2, device=dev, dtype=torch.int32)
        rm.march_rays_train(rays_o, rays_d, self.bitfield, c.bound, c.dt_gamma, c.max_steps, N, c.cascade,
                            c.grid_size, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
        # ---- hash grid ----
        L, C = c.num_levels, c.level_dim
        x01 = (xyzs + c.bound) / (2 * c.bound)
        emb_h = self.embeddings.to(f16)
        enc_lbc = torch.empty(L, M, C, device=dev, dtype=f16)
        ge.grid_encode_forward(x01, emb_h, self.offsets, enc_lbc, M, 3, C, L, self.S, c.base_resolution, None, 0,
                               False, 0)
        enc = enc_lbc.permute(1, 0, 2).reshape(M, L * C)
        # ---- density MLP ----
        ws_h = self.w_sigma.to(f16)
        fb_s = torch.empty(c.sigma_layers, M, c.hidden_dim, device=dev, dtype=f16)
        sig_out = torch.empty(M, 16, device=dev, dtype=f16)
        ff.ffmlp_forward(enc, ws_h, M, self.enc_dim, 16, c.hidden_dim, c.sigma_layers, 0, 6, fb_s, sig_out)
        h0 = sig_out[:, 0].float()
        sigma = torch.exp(h0) * c.density_scale
        # ---- LiDAR head ----
        fenc = torch.empty(M, self.nfreq, device=dev, dtype=f32)
        fe.freq_encode_forward(dirs, M, 3, c.freq_degree, self.nfreq, fenc)
        head_in = torch.zeros(M, c.head_in_dim, device=dev, dtype=f16)
        head_in[:, :self.nfreq] = fenc.to(f16)
        head_in[:, self.nfreq:self.nfreq + 15] = sig_out[:, 1:16]
        wh_h = self.w_head.to(f16)
        fb_h = torch.empty(c.head_layers, M, c.hidden_dim, device=dev, dtype=f16)
        head_out = torch.empty(M, 16, device=dev, dtype=f16)
        ff.ffmlp_forward(head_in, wh_h, M, c.head_in_dim, 16, c.hidden_dim, c.head_layers, 0, 6, fb_h, head_out)
        rgb3 = torch.zeros(M, 3, device=dev, dtype=f32)
        rgb3[:, :2] = torch.sigmoid(head_out[:, :2].float())
        # ---- composite ----
        wsum = torch.empty(N, device=dev, dtype=f32)
        depth = torch.empty(N, device=dev, dtype=f32)
        image = torch.empty(N, 3, device=dev, dtype=f32)
        rm.composite_rays_train_forward(sigma, rgb3, deltas, rays, M, N, c.T_thresh, wsum, depth, image)
        # ---- loss (nerf/utils.py:726-734) ----
        dt_min = 2 * 3 ** 0.5 / c.max_steps
        dt_max = 2 * 3 ** 0.5 * (1 << (c.cascade - 1)) / c.grid_size
        t0 = nears + (nears * c.dt_gamma).clamp(dt_min, dt_max) * noises
        m = gt[:, 0]
        gi, gd = gt[:, 1] * m, gt[:, 2] * m
        D = depth + t0 * wsum
        e_d, e_r, e_i = D * m - gd, image[:, 0] - m, image[:, 1] * m - gi
        loss = (c.alpha_d * e_d.abs() + c.alpha_r * e_r ** 2 + c.alpha_i * e_i ** 2).mean()
        s = c.loss_scale / N
        gD = c.alpha_d * m * torch.sign(e_d) * s
        g_ws = (gD * t0).contiguous()
        g_img = torch.zeros(N, 3, device=dev, dtype=f32)
        g_img[:, 0] = 2 * c.alpha_r * e_r * s
        g_img[:, 1] = 2 * c.alpha_i * e_i * m * s
        # ---- backward ----
        g_sigma = torch.zeros(M, device=dev, dtype=f32)
        g_rgb = torch.zeros(M, 3, device=dev, dtype=f32)
        rm.composite_rays_train_backward(g_ws, g_img, sigma, rgb3, deltas, rays, wsum, image, M, N, c.T_thresh,
                                         g_sigma, g_rgb)
        g_head_out = torch.zeros(M, 16, device=dev, dtype=f16)
        g_head_out[:, :2] = (g_rgb[:, :2] * rgb3[:, :2] * (1 - rgb3[:, :2])).to(f16)
        bb_h = torch.zeros(c.head_layers, M, c.hidden_dim, device=dev, dtype=f16)
        g_head_in = torch.zeros(M, c.head_in_dim, device=dev, dtype=f16)
        gw_head = torch.zeros_like(wh_h)
        ff.ffmlp_backward(g_head_out, head_in, wh_h, fb_h, M, c.head_in_dim, 16, c.hidden_dim, c.head_layers, 0, 6,
                          True, bb_h, g_head_in, gw_head)
        g_sig_out = torch.zeros(M, 16, device=dev, dtype=f16)
        g_sig_out[:, 0] = (g_sigma * c.density_scale * torch.exp(h0.clamp(-15, 15))).to(f16)
        g_sig_out[:, 1:16] = g_head_in[:, self.nfreq:self.nfreq + 15]
        bb_s = torch.zeros(c.sigma_layers, M, c.hidden_dim, device=dev, dtype=f16)
        g_enc = torch.zeros(M, self.enc_dim, device=dev, dtype=f16)
        gw_sigma = torch.zeros_like(ws_h)
        ff.ffmlp_backward(g_sig_out, enc, ws_h, fb_s, M, self.enc_dim, 16, c.hidden_dim, c.sigma_layers, 0, 6, True,
                          bb_s, g_enc, gw_sigma)
        g_lbc = g_enc.view(M, L, C).permute(1, 0, 2).contiguous()
        g_emb = torch.zeros_like(emb_h)
        ge.grid_encode_backward(g_lbc, x01, emb_h, self.offsets, g_emb, M, 3, C, L, self.S, c.base_resolution, None,
                                None, 0, False, 0)
        if apply_adam:
            inv = 1.0 / c.loss_scale
            self.embeddings.grad.copy_(g_emb.float() * inv)
            self.w_sigma.grad.copy_(gw_sigma.float() * inv)
            self.w_head.grad.copy_(gw_head.float() * inv)
            self.opt.step()
        return dict(loss=loss, counter=counter, g_emb=g_emb, gw_sigma=gw_sigma, gw_head=gw_head, depth=depth,
                    image=image, wsum=wsum)
