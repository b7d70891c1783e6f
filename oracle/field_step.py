"""CPU restatement of ONE training step of the LiDAR field in occupancy-march mode, composed from the oracle's
per-op restatements (oracle/lnb_oracle.c).  TEST INFRASTRUCTURE: used by tests/ (engine parity), smoke() and
bench.py's cpu_baseline / --impl reference legs only.

Graph (reference file:line in each op's C restatement):
  march_rays_train -> grid_encode(fp16 table) -> FFMLP(32->64->64->16) -> sigma = exp(h0), geo = h[1:16]
  -> [freq_encode(dir, 12) | geo | 0] (96) -> FFMLP(96->64->64->16) -> sigmoid(h[0:2])
  -> composite_rays_train (2 channels) -> LiDAR loss (nerf/utils.py:726-734) -> backward of all of it -> Adam.
"""
import numpy as np

from . import oracle as orc


class FieldParams:
    """Flat fp32 parameter vector [table | sigma MLP | head MLP] + Adam state, initialised like the engine."""

    def __init__(self, cfg, P=None):
        self.cfg = cfg
        c = cfg
        self.pls = float(np.exp2(np.log2(c.desired_resolution / c.base_resolution) / (c.num_levels - 1)))
        self.offsets = orc.grid_offsets(3, c.num_levels, c.base_resolution, self.pls, c.log2_hashmap_size, False)
        self.n_rows = int(self.offsets[-1])
        self.enc_dim = c.num_levels * c.level_dim
        self.n_table = self.n_rows * c.level_dim
        self.n_sigma = c.hidden_dim * (self.enc_dim + c.hidden_dim * (c.sigma_layers - 1) + 16)
        self.n_head = c.hidden_dim * (c.head_in_dim + c.hidden_dim * (c.head_layers - 1) + 16)
        n = self.n_table + self.n_sigma + self.n_head
        self.P = np.zeros(n, np.float32) if P is None else np.ascontiguousarray(P, np.float32).copy()
        self.m = np.zeros(n, np.float32)
        self.v = np.zeros(n, np.float32)
        self.step = 0

    def split(self, flat):
        a, b = self.n_table, self.n_table + self.n_sigma
        return flat[:a].reshape(self.n_rows, self.cfg.level_dim), flat[a:b], flat[b:]


def lidar_loss(D, image, gt, a_d, a_r, a_i, loss_scale=1.0, patch=(1, 1), a_grad=0.0, inv_scale=1.0, clip=0.01):
    """LiDAR loss of Trainer.train_step (nerf/utils.py:707-734) and its gradient w.r.t. the rendered absolute depth D [N]
    and image [N,2]; gt rows = (ray-drop mask, intensity, depth).  With `patch` = (px, py) > 1 and a_grad > 0 it includes
    the patch depth-gradient term (:748-876, sobel_grad = False, depth_grad_loss = l1): the rays are patches
    [N / (px py), px, py] and the HORIZONTAL differences of P = D m / scale are compared - in absolute value - with the
    signed differences of G = d_gt m / scale where the ground truth is locally flat (|dG| < clip) and returned.
    Returns (loss, g_D * loss_scale, g_image * loss_scale)."""
    N = D.shape[0]
    m = gt[:, 0]
    gi, gd = gt[:, 1] * m, gt[:, 2] * m
    e_d, e_r, e_i = D * m - gd, image[:, 0] - m, image[:, 1] * m - gi
    loss = float(np.mean(a_d * np.abs(e_d) + a_r * e_r ** 2 + a_i * e_i ** 2))
    s = np.float32(loss_scale / N)
    gD = (a_d * m * np.sign(e_d) * s).astype(np.float32)
    g_img = np.stack([2 * a_r * e_r * s, 2 * a_i * e_i * m * s], -1).astype(np.float32)
    px, py = patch
    if a_grad > 0 and px * py > 1 and py > 1:
        n_patch = N // (px * py)
        n = n_patch * px * py
        P = (D[:n] * m[:n] * np.float32(inv_scale)).reshape(n_patch, px, py)
        G = (gd[:n] * np.float32(inv_scale)).reshape(n_patch, px, py)
        mm = m[:n].reshape(n_patch, px, py)
        dps = P[:, :, :-1] - P[:, :, 1:]
        dg = G[:, :, :-1] - G[:, :, 1:]
        msk = mm[:, :, :-1] * (np.abs(dg) < clip)
        e = np.abs(dps) * msk - dg * msk
        loss += float(a_grad * np.mean(np.abs(e)))
        contrib = (a_grad / e.size) * np.sign(e) * msk * np.sign(dps)
        gP = np.zeros_like(P)
        gP[:, :, :-1] += contrib
        gP[:, :, 1:] -= contrib
        gD[:n] += (gP.reshape(-1) * m[:n] * np.float32(inv_scale) * np.float32(loss_scale)).astype(np.float32)
    return loss, gD, g_img


def field_step(params: FieldParams, rays_o, rays_d, gt, noises, bitfield, M, level_scales=None, apply_adam=True):
    """Returns dict(loss, grad (flat fp32, unscaled), counts, ws, depth, image, n_samples)."""
    c = params.cfg
    N = rays_o.shape[0]
    bf16 = getattr(c, "mlp_dtype", "fp16") == "bf16"
    orc.set_mlp_dtype("bf16" if bf16 else "fp16")
    Ph = orc.to_half(params.P)
    table, w_sigma, w_head = params.split(Ph)
    if bf16:      # the MLP weights are the bf16 rounding of the fp32 master (the table stays fp16)
        _, w_sigma, w_head = params.split(orc.to_bf16(params.P))
    nears = np.full(N, c.min_near_lidar, np.float32)
    fars = nears * np.float32(c.far_factor)
    xyzs, dirs, deltas, rays, counter = orc.march_rays_train(rays_o, rays_d, c.bound, bitfield, c.cascade, c.grid_size,
                                                             nears, fars, noises, c.dt_gamma, c.max_steps, M)
    dt_min = np.float32(2 * 1.7320508075688772) / np.float32(c.max_steps)
    dt_max = np.float32(2 * 1.7320508075688772) * np.float32(1 << (c.cascade - 1)) / np.float32(c.grid_size)
    t0 = nears + np.clip(nears * np.float32(c.dt_gamma), dt_min, dt_max) * noises

    x01 = ((xyzs + np.float32(c.bound)) * np.float32(1.0 / (2.0 * c.bound))).astype(np.float32)
    enc = orc.grid_encode_forward(x01, table, params.offsets, params.pls, c.base_resolution, 0, False, 0, True, False,
                                  level_scales)
    enc = orc.mlp_round(enc)       # (fp16 interpolation result -> the MLP's operand type)
    sig_out, fb_s = orc.ffmlp_forward(enc, w_sigma, params.enc_dim, 16, c.hidden_dim, c.sigma_layers)
    sigma = np.exp(sig_out[:, 0]).astype(np.float32) * np.float32(c.density_scale)
    # direction encoding of the head: frequency (network.py:83) or spherical harmonics (network.py:64), cfg.dir_encoding
    if getattr(c, "dir_encoding", "frequency") == "sh":
        fenc = orc.sh_encode_forward(dirs, c.sh_degree)
    else:
        fenc = orc.freq_encode_forward(dirs, c.freq_degree)
    head_in = np.zeros((M, c.head_in_dim), np.float32)
    head_in[:, :fenc.shape[1]] = fenc
    head_in[:, fenc.shape[1]:fenc.shape[1] + 15] = sig_out[:, 1:16]
    head_in = orc.mlp_round(head_in)
    head_out, fb_h = orc.ffmlp_forward(head_in, w_head, c.head_in_dim, 16, c.hidden_dim, c.head_layers)
    rgb = (1.0 / (1.0 + np.exp(-head_out[:, :2]))).astype(np.float32)
    ws, depth, image = orc.composite_rays_train_forward(sigma, rgb, deltas, rays, c.T_thresh)

    # loss (nerf/utils.py:726-734, + the patch term of :748-876) with absolute depth = depth + t0 * ws
    D = depth + t0 * ws
    px, py = getattr(c, "patch_size", (1, 1))
    a_grad = getattr(c, "alpha_grad", 0.0) if px * py > 1 else 0.0
    loss, gD, g_img = lidar_loss(D, image, gt, c.alpha_d, c.alpha_r, c.alpha_i, c.loss_scale, (px, py), a_grad,
                                 1.0 / c.min_near_lidar, getattr(c, "grad_clip", 0.01))
    g_ws = (gD * t0).astype(np.float32)

    g_sigma, g_rgb = orc.composite_rays_train_backward(g_ws, g_img, sigma, rgb, deltas, rays, ws, image, c.T_thresh, gD,
                                                       depth)
    g_head_out = np.zeros((M, 16), np.float32)
    g_head_out[:, :2] = g_rgb * rgb * (1 - rgb)
    g_head_in, gw_head, _ = orc.ffmlp_backward(g_head_out, head_in, w_head, fb_h, c.head_in_dim, 16, c.hidden_dim,
                                               c.head_layers, True)
    g_sig_out = np.zeros((M, 16), np.float32)
    g_sig_out[:, 0] = g_sigma * np.float32(c.density_scale) * np.exp(np.clip(sig_out[:, 0], -15, 15))
    g_sig_out[:, 1:16] = g_head_in[:, fenc.shape[1]:fenc.shape[1] + 15]
    g_enc, gw_sigma, _ = orc.ffmlp_backward(g_sig_out, enc, w_sigma, fb_s, params.enc_dim, 16, c.hidden_dim,
                                            c.sigma_layers, True, grad_inputs_fp16=True)
    g_table = orc.grid_encode_backward(orc.to_half(g_enc), x01, table.shape, params.offsets, params.pls, c.base_resolution,
                                       0, False, 0, False, None, level_scales)
    grad = np.concatenate([g_table.reshape(-1), gw_sigma, gw_head]).astype(np.float32)
    if apply_adam:
        params.step += 1
        orc.adam_step(params.P, grad, params.m, params.v, c.lr, c.beta1, c.beta2, c.eps, params.step, 1.0 / c.loss_scale)
    orc.set_mlp_dtype("fp16")
    return dict(loss=loss, grad=grad, counts=rays[:, 2].copy(), ws=ws, depth=depth, image=image,
                n_samples=int(counter[0]), sigma=sigma, rgb=rgb, enc=enc, sig_out=sig_out, head_out=head_out)
