"""numpy front-end of the CPU oracle (oracle/lnb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this
module; the product package (lidar-nerf_b200/) never does and has no CPU fallback.

Each function mirrors the argument meaning of the reference binding it restates (see lnb_oracle.c for the
reference file:line citations) but takes/returns numpy arrays.  Arrays the reference stores as fp16 are
float32 arrays holding fp16-representable values here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liblnb_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/lnb_oracle.c with the system gcc (OpenMP if available)."""
    src = os.path.join(_HERE, "lnb_oracle.c")
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    base = ["-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-shared"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            r = subprocess.run([cc, *base, *omp, "-o", _LIB_PATH, src, "-lm"], capture_output=True, text=True)
            if r.returncode == 0:
                return _LIB_PATH
            last = r.stderr
    raise RuntimeError("could not compile the CPU oracle:\n" + str(last))


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    lib().orc_set_threads(C.c_int(int(n)))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


u32, f32, i32 = C.c_uint32, C.c_float, C.c_int


# ----------------------------------------------------------------------------- raymarching
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    o, d, bb = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3), _f(aabb)
    n = o.shape[0]
    nears, fars = np.empty(n, np.float32), np.empty(n, np.float32)
    lib().orc_near_far_from_aabb(_p(o), _p(d), _p(bb), u32(n), f32(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    o, d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    out = np.empty((o.shape[0], 2), np.float32)
    lib().orc_sph_from_ray(_p(o), _p(d), f32(radius), u32(o.shape[0]), _p(out))
    return out


def morton3D(coords):
    c = _i(coords).reshape(-1, 3)
    out = np.empty(c.shape[0], np.int32)
    lib().orc_morton3D(_p(c), u32(c.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    x = _i(indices).reshape(-1)
    out = np.empty((x.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(x), u32(x.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    g = _f(grid).reshape(-1)
    n = g.shape[0] // 8
    out = np.empty(n, np.uint8)
    lib().orc_packbits(_p(g), u32(n), f32(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, cascade, H, nears, fars, noises, dt_gamma=0.0,
                     max_steps=1024, M=None):
    """Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3] (ray order), counter [2]."""
    o, d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    n = o.shape[0]
    M = n * max_steps if M is None else int(M)
    xyzs, dirs = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    deltas, rays = np.zeros((M, 2), np.float32), np.zeros((n, 3), np.int32)
    counter = np.zeros(2, np.int32)
    bf = np.ascontiguousarray(bitfield, dtype=np.uint8)
    lib().orc_march_rays_train(_p(o), _p(d), _p(bf), f32(bound), f32(dt_gamma), u32(max_steps), u32(n), u32(cascade),
                               u32(H), u32(M), _p(_f(nears)), _p(_f(fars)), _p(xyzs), _p(dirs), _p(deltas), _p(rays),
                               _p(counter), _p(_f(noises)))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, cascade, H, nears, fars, noises,
               dt_gamma=0.0, max_steps=1024):
    o, d = _f(rays_o).reshape(-1, 3), _f(rays_d).reshape(-1, 3)
    M = n_alive * n_step
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    bf = np.ascontiguousarray(bitfield, dtype=np.uint8)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(_i(rays_alive)), _p(_f(rays_t)), _p(o), _p(d), f32(bound),
                         f32(dt_gamma), u32(max_steps), u32(cascade), u32(H), _p(bf), _p(_f(nears)), _p(_f(fars)),
                         _p(xyzs), _p(dirs), _p(deltas), _p(_f(noises)))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    s, c, dl, r = _f(sigmas).reshape(-1), _f(rgbs), _f(deltas), _i(rays)
    ch = c.shape[1]
    N = r.shape[0]
    ws, dep, img = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, ch), np.float32)
    lib().orc_composite_rays_train_forward(_p(s), _p(c), _p(dl), _p(r), u32(s.shape[0]), u32(N), f32(T_thresh), u32(ch),
                                           _p(ws), _p(dep), _p(img))
    return ws, dep, img


def composite_rays_train_backward(g_ws, g_img, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4,
                                  g_depth=None, depth=None):
    s, c, dl, r = _f(sigmas).reshape(-1), _f(rgbs), _f(deltas), _i(rays)
    ch = c.shape[1]
    gs, gc = np.zeros_like(s), np.zeros_like(c)
    gd = None if g_depth is None else _f(g_depth)
    dp = None if depth is None else _f(depth)
    lib().orc_composite_rays_train_backward(_p(_f(g_ws)), _p(gd), _p(_f(g_img)), _p(s), _p(c), _p(dl), _p(r),
                                            _p(_f(weights_sum)), _p(dp), _p(_f(image)), u32(s.shape[0]),
                                            u32(r.shape[0]), f32(T_thresh), u32(ch), _p(gs), _p(gc))
    return gs, gc


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive (int32), rays_t, weights_sum, depth, image (float32, C-contiguous)."""
    lib().orc_composite_rays(u32(n_alive), u32(n_step), f32(T_thresh), _p(rays_alive), _p(rays_t), _p(_f(sigmas)),
                             _p(_f(rgbs)), _p(_f(deltas)), _p(weights_sum), _p(depth), _p(image))


# ----------------------------------------------------------------------------- encoders
def grid_offsets(input_dim=3, num_levels=16, base_resolution=16, per_level_scale=2.0, log2_hashmap_size=19,
                 align_corners=False):
    """Level table of the reference's GridEncoder.__init__ (gridencoder/grid.py:179-192)."""
    offsets, offset = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(2 ** log2_hashmap_size, (res if align_corners else res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def grid_encode_forward(inputs, table, offsets, per_level_scale, base_resolution, gridtype=0, align_corners=False,
                        interp=0, half=False, calc_grad_inputs=False, level_scales=None):
    """-> outputs [B, L*C] (and dy_dx [B, L*D*C] if requested)."""
    x, t, off = _f(inputs), _f(table), _i(offsets)
    B, D = x.shape
    Cc, L = t.shape[1], off.shape[0] - 1
    out = np.empty((B, L * Cc), np.float32)
    dy = np.empty((B, L * D * Cc), np.float32) if calc_grad_inputs else None
    S = np.float32(np.log2(per_level_scale))
    lib().orc_grid_encode_forward(_p(x), _p(t), _p(off), _p(out), u32(B), u32(D), u32(Cc), u32(L), f32(S),
                                  u32(base_resolution), _p(dy), u32(gridtype), i32(int(align_corners)), u32(interp),
                                  i32(int(half)), i32(1), _p(None if level_scales is None else _f(level_scales)))
    return (out, dy) if calc_grad_inputs else out


def grid_encode_backward(grad, inputs, table_shape, offsets, per_level_scale, base_resolution, gridtype=0,
                         align_corners=False, interp=0, half=False, dy_dx=None, level_scales=None):
    g, x, off = _f(grad), _f(inputs), _i(offsets)
    B, D = x.shape
    Cc, L = table_shape[1], off.shape[0] - 1
    gt = np.zeros(table_shape, np.float32)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    S = np.float32(np.log2(per_level_scale))
    lib().orc_grid_encode_backward(_p(g), _p(x), _p(off), _p(gt), u32(B), u32(D), u32(Cc), u32(L), f32(S),
                                   u32(base_resolution), _p(None if dy_dx is None else _f(dy_dx)), _p(gi), u32(gridtype),
                                   i32(int(align_corners)), u32(interp), i32(int(half)), i32(1),
                                   _p(None if level_scales is None else _f(level_scales)))
    return (gt, gi) if dy_dx is not None else gt


def freq_encode_forward(inputs, degree):
    x = _f(inputs)
    B, D = x.shape
    Cc = D + 2 * D * degree
    out = np.empty((B, Cc), np.float32)
    lib().orc_freq_encode_forward(_p(x), u32(B), u32(D), u32(degree), u32(Cc), _p(out))
    return out


def freq_encode_backward(grad, outputs, D, degree):
    g, o = _f(grad), _f(outputs)
    B, Cc = g.shape
    gi = np.empty((B, D), np.float32)
    lib().orc_freq_encode_backward(_p(g), _p(o), u32(B), u32(D), u32(degree), u32(Cc), _p(gi))
    return gi


def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    x = _f(inputs)
    B = x.shape[0]
    out = np.empty((B, degree * degree), np.float32)
    dy = np.empty((B, 3 * degree * degree), np.float32) if calc_grad_inputs else None
    lib().orc_sh_encode_forward(_p(x), _p(out), u32(B), u32(degree), _p(dy))
    return (out, dy) if calc_grad_inputs else out


def sh_encode_backward(grad, degree, dy_dx):
    g, dy = _f(grad), _f(dy_dx)
    gi = np.zeros((g.shape[0], 3), np.float32)
    lib().orc_sh_encode_backward(_p(g), u32(g.shape[0]), u32(degree), _p(dy), _p(gi))
    return gi


# ----------------------------------------------------------------------------- ffmlp
def to_half(a):
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def to_bf16(a):
    """Round to bfloat16 (nearest even), returned as float32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


_MLP_BF16 = False


def set_mlp_dtype(name):
    """Element type of the fused MLPs in ffmlp_forward / ffmlp_backward: "fp16" (default) or "bf16" (the -DLNB_BF16
    build of the kernels)."""
    global _MLP_BF16
    _MLP_BF16 = name == "bf16"
    lib().orc_set_mlp_mode(C.c_int(1 if _MLP_BF16 else 0))


def mlp_round(a):
    return to_bf16(a) if _MLP_BF16 else to_half(a)


def ffmlp_forward(inputs, weights, input_dim, output_dim, hidden_dim, num_layers):
    """inputs [B,in], weights flat; both are rounded to fp16 first.  -> outputs [B,out], forward_buffer."""
    x, w = mlp_round(inputs), mlp_round(weights).reshape(-1)
    B = x.shape[0]
    fb = np.empty((num_layers, B, hidden_dim), np.float32)
    out = np.empty((B, output_dim), np.float32)
    lib().orc_ffmlp_forward(_p(_f(x)), _p(_f(w)), u32(B), u32(input_dim), u32(output_dim), u32(hidden_dim),
                            u32(num_layers), _p(fb), _p(out))
    return out, fb


def ffmlp_backward(grad, inputs, weights, forward_buffer, input_dim, output_dim, hidden_dim, num_layers,
                   calc_grad_inputs=True, grad_inputs_fp16=False):
    """grad_inputs_fp16: in bf16 mode, round the input gradient to fp16 nevertheless (density MLP -> hash-grid scatter)."""
    g, x, w = mlp_round(grad), mlp_round(inputs), mlp_round(weights).reshape(-1)
    if _MLP_BF16:
        lib().orc_set_mlp_mode(C.c_int(3 if grad_inputs_fp16 else 1))
    B = x.shape[0]
    bb = np.zeros((num_layers, B, hidden_dim), np.float32)
    gi = np.zeros((B, input_dim), np.float32) if calc_grad_inputs else None
    gw = np.zeros_like(w)
    lib().orc_ffmlp_backward(_p(_f(g)), _p(_f(x)), _p(_f(w)), _p(_f(forward_buffer)), u32(B), u32(input_dim),
                             u32(output_dim), u32(hidden_dim), u32(num_layers), _p(bb), _p(gi), _p(gw))
    return gi, gw, bb


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """In place on p, m, v (float32 contiguous)."""
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    lib().orc_adam_step(_p(p), _p(_f(g)), _p(m), _p(v), C.c_size_t(p.size), f32(lr), f32(beta1), f32(beta2), f32(eps),
                        f32(bc1), f32(bc2), f32(grad_scale))


# ----------------------------------------------------------------------------- activations (activation.py:6-20)
def trunc_exp_forward(x):
    return np.exp(np.asarray(x, np.float32))


def trunc_exp_backward(g, x):
    return np.asarray(g, np.float32) * np.exp(np.clip(np.asarray(x, np.float32), -15, 15))


def grad_total_variation(inputs, table, grad, offsets, weight, per_level_scale, base_resolution, gridtype=0,
                         align_corners=False, level_scales=None):
    """gridencoder.cu:695-808 (fp32): returns grad + TV gradient."""
    x, t, off = _f(inputs), _f(table), _i(offsets)
    g = _f(grad).copy()
    B, D = x.shape
    S = np.float32(np.log2(per_level_scale))
    lib().orc_grad_total_variation(_p(x), _p(t), _p(g), _p(off), f32(weight), u32(B), u32(D), u32(t.shape[1]),
                                   u32(off.shape[0] - 1), f32(S), u32(base_resolution), u32(gridtype),
                                   i32(int(align_corners)), _p(None if level_scales is None else _f(level_scales)))
    return g


def chamfer_forward(xyz1, xyz2):
    """extern/chamfer3D/chamfer3D.cu:135-166: (dist1 [B,N], dist2 [B,M], idx1, idx2)."""
    a, b = _f(xyz1), _f(xyz2)
    Bn, N, _ = a.shape
    M = b.shape[1]
    d1, i1 = np.empty((Bn, N), np.float32), np.empty((Bn, N), np.int32)
    d2, i2 = np.empty((Bn, M), np.float32), np.empty((Bn, M), np.int32)
    lib().orc_chamfer_nn(_p(a), _p(b), u32(Bn), u32(N), u32(M), _p(d1), _p(i1))
    lib().orc_chamfer_nn(_p(b), _p(a), u32(Bn), u32(M), u32(N), _p(d2), _p(i2))
    return d1, d2, i1, i2


def lidar_to_pano_with_intensities(points, H, W, lidar_K, max_depth=80):
    """convert.py:99-160: points [N,4] (or [N,3]) -> (pano [H,W], intensities [H,W])."""
    p = _f(points)
    pano, inten = np.empty((H, W), np.float32), np.empty((H, W), np.float32)
    lib().orc_lidar_to_pano(_p(p), u32(p.shape[1]), u32(p.shape[0]), u32(H), u32(W), f32(lidar_K[0]), f32(lidar_K[1]),
                            f32(max_depth), _p(pano), _p(inten))
    return pano, inten


def pano_to_lidar_with_intensities(pano, intensities, lidar_K):
    """convert.py:194-235: -> points [n,4] in row-major order of the non-empty pixels."""
    pa = _f(pano)
    H, W = pa.shape
    out = np.empty((H * W, 4), np.float32)
    lib().orc_pano_to_lidar.restype = C.c_uint32
    n = lib().orc_pano_to_lidar(_p(pa), _p(None if intensities is None else _f(intensities)), u32(H), u32(W),
                                f32(lidar_K[0]), f32(lidar_K[1]), _p(out))
    return out[:int(n)].copy()
