#!/usr/bin/env python3
"""Freeze outputs of the reference's OWN CUDA kernels into fixtures.  Run ON THE GPU BOX:

    gpurun -- python tests/golden/make_golden_gpu.py gpurun_out/golden

with oracle/_ref/ built by oracle/build_ref.py (the unmodified reference extensions for sm_100a).  The npz
files it writes are copied into tests/golden/ref_cuda_*.npz and committed; they pin the CPU oracle
(tests/test_oracle_goldens.py, no GPU needed) and are re-checked against our kernels (tests -m gpu).

Inputs come from tests/cases.py with fixed seeds, so the tests can regenerate them instead of storing them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
import refcuda  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_gpu_out")
os.makedirs(out_dir, exist_ok=True)
dev = torch.device("cuda:0")


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t if dtype is None else t.to(dtype)


def save(name, **arrs):
    np.savez_compressed(os.path.join(out_dir, name), **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                        for k, v in arrs.items()})
    print("wrote", name)


MARCH_CASES = [
    dict(seed=11, N=256, cascade=1, bound=1.0, fill=0.3, lidar=False, dt_gamma=0.0, max_steps=1024, near_far="aabb"),
    dict(seed=12, N=256, cascade=1, bound=1.0, fill=0.15, lidar=True, dt_gamma=0.0, max_steps=1024, near_far="lidar"),
    dict(seed=13, N=256, cascade=3, bound=4.0, fill=0.3, lidar=False, dt_gamma=1.0 / 128, max_steps=128, near_far="aabb"),
    dict(seed=14, N=128, cascade=1, bound=1.0, fill=0.9, lidar=False, dt_gamma=1.0 / 128, max_steps=64, near_far="aabb"),
]
COMPOSITE_CASES = [dict(seed=21, N=300, max_count=200), dict(seed=22, N=64, max_count=700, opaque_frac=0.8)]
GRID_CASES = [
    dict(seed=31, B=1000, D=3, C=2, L=16, desired_resolution=2048, half=False),
    dict(seed=32, B=1000, D=3, C=2, L=16, desired_resolution=32768, half=True),
    dict(seed=33, B=500, D=2, C=4, L=4, desired_resolution=2048, half=False),
    dict(seed=34, B=500, D=3, C=1, L=8, desired_resolution=512, half=False),
]


def main():
    rm = refcuda.load("_raymarching")
    if rm is not None:
        # near/far, morton, packbits
        rng = np.random.default_rng(1)
        o = rng.uniform(-1.5, 1.5, size=(400, 3)).astype(np.float32)
        d = cases.unit(rng.normal(size=(400, 3))).astype(np.float32)
        aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
        nears, fars = torch.empty(400, device=dev), torch.empty(400, device=dev)
        rm.near_far_from_aabb(T(o), T(d), T(aabb), 400, 0.05, nears, fars)
        coords = torch.empty(400, 2, device=dev)
        rm.sph_from_ray(T(o * 0.3), T(d), 2.0, 400, coords)
        c3 = rng.integers(0, 128, size=(500, 3)).astype(np.int32)
        idx = torch.empty(500, dtype=torch.int32, device=dev)
        rm.morton3D(T(c3), 500, idx)
        back = torch.empty(500, 3, dtype=torch.int32, device=dev)
        rm.morton3D_invert(idx, 500, back)
        grid = rng.uniform(0, 1, size=(1, 4096 * 8)).astype(np.float32)
        bits = torch.empty(4096, dtype=torch.uint8, device=dev)
        rm.packbits(T(grid), 4096, 0.5, bits)
        save("ref_cuda_utils.npz", o=o, d=d, aabb=aabb, min_near=0.05, nears=nears, fars=fars, sph=coords, radius=2.0,
             coords=c3, morton=idx, morton_inv=back, grid=grid, thresh=0.5, bits=bits)

        # march_rays_train: per-ray results in RAY ORDER (arrival order of the atomics is not deterministic)
        for i, cs in enumerate(MARCH_CASES):
            c = cases.march_case(cs["seed"], cs["N"], cs["cascade"], cs["bound"], 128, cs["fill"], cs["lidar"])
            N = cs["N"]
            ro, rd = T(c["rays_o"]), T(c["rays_d"])
            if cs["near_far"] == "aabb":
                b = cs["bound"]
                nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
                rm.near_far_from_aabb(ro, rd, T(np.array([-b, -b, -b, b, b, b], np.float32)), N, 0.05, nears, fars)
            else:
                nears = torch.full((N,), 0.0108, device=dev)
                fars = nears * 81.0
            M = N * cs["max_steps"]
            xyzs, dirs = torch.zeros(M, 3, device=dev), torch.zeros(M, 3, device=dev)
            deltas = torch.zeros(M, 2, device=dev)
            rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
            counter = torch.zeros(2, dtype=torch.int32, device=dev)
            rm.march_rays_train(ro, rd, T(c["bitfield"]), cs["bound"], cs["dt_gamma"], cs["max_steps"], N, cs["cascade"],
                                128, M, nears, fars, xyzs, dirs, deltas, rays, counter, T(c["noises"]))
            torch.cuda.synchronize()
            rays_np = rays.cpu().numpy()
            order = np.argsort(rays_np[:, 0])
            rays_np = rays_np[order]
            counts = rays_np[:, 2]
            # re-pack samples in ray order
            xs, ds = xyzs.cpu().numpy(), deltas.cpu().numpy()
            pack_x = np.concatenate([xs[o_:o_ + n_] for _, o_, n_ in rays_np]) if counts.sum() else np.zeros((0, 3), np.float32)
            pack_d = np.concatenate([ds[o_:o_ + n_] for _, o_, n_ in rays_np]) if counts.sum() else np.zeros((0, 2), np.float32)
            save(f"ref_cuda_march{i}.npz", cfg=np.array(repr(cs)), nears=nears, fars=fars, counts=counts, xyzs=pack_x,
                 deltas=pack_d, total=counter.cpu().numpy())

        # composite fwd/bwd
        for i, cs in enumerate(COMPOSITE_CASES):
            c = cases.composite_case(**cs)
            N, M = c["N"], c["M"]
            ws, dep, img = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
            sig, rgb, dl, rays = T(c["sigmas"]), T(c["rgbs"]), T(c["deltas"]), T(c["rays"])
            rm.composite_rays_train_forward(sig, rgb, dl, rays, M, N, 1e-4, ws, dep, img)
            rng = np.random.default_rng(cs["seed"] + 100)
            gws, gimg = rng.normal(size=N).astype(np.float32), rng.normal(size=(N, 3)).astype(np.float32)
            gs, gc = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
            rm.composite_rays_train_backward(T(gws), T(gimg), sig, rgb, dl, rays, ws, img, M, N, 1e-4, gs, gc)
            save(f"ref_cuda_composite{i}.npz", cfg=np.array(repr(cs)), weights_sum=ws, depth=dep, image=img, g_ws=gws,
                 g_img=gimg, grad_sigmas=gs, grad_rgbs=gc)

        # inference march + composite (one round)
        c = cases.march_case(15, 128, 1, 1.0, 128, 0.3, False)
        N, n_step = 128, 8
        ro, rd = T(c["rays_o"]), T(c["rays_d"])
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        rm.near_far_from_aabb(ro, rd, T(np.array([-1, -1, -1, 1, 1, 1], np.float32)), N, 0.05, nears, fars)
        alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        xyzs, dirs, deltas = torch.zeros(N * n_step, 3, device=dev), torch.zeros(N * n_step, 3, device=dev), torch.zeros(N * n_step, 2, device=dev)
        rm.march_rays(N, n_step, alive, rays_t, ro, rd, 1.0, 0.0, 1024, 1, 128, T(c["bitfield"]), nears, fars, xyzs, dirs,
                      deltas, T(c["noises"]))
        rng = np.random.default_rng(16)
        sig = rng.gamma(1.0, 20.0, size=N * n_step).astype(np.float32)
        rgb = rng.uniform(0, 1, size=(N * n_step, 3)).astype(np.float32)
        ws, dep, img = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
        alive2, rays_t2 = alive.clone(), rays_t.clone()
        rm.composite_rays(N, n_step, 1e-2, alive2, rays_t2, T(sig), T(rgb), deltas, ws, dep, img)
        save("ref_cuda_infer.npz", nears=nears, fars=fars, xyzs=xyzs, deltas=deltas, sigmas=sig, rgbs=rgb, alive=alive2,
             rays_t=rays_t2, weights_sum=ws, depth=dep, image=img)

    ge = refcuda.load("_gridencoder")
    if ge is not None:
        for i, cs in enumerate(GRID_CASES):
            c = cases.grid_case(cs["seed"], cs["B"], cs["D"], cs["C"], cs["L"], 16, cs["desired_resolution"])
            dt = torch.half if cs["half"] else torch.float32
            B, D, C, L = cs["B"], cs["D"], cs["C"], cs["L"]
            S = float(np.log2(c["per_level_scale"]))
            emb = T(c["table"]).to(dt)
            out = torch.empty(L, B, C, device=dev, dtype=dt)
            dy = torch.empty(B, L * D * C, device=dev, dtype=dt)
            ge.grid_encode_forward(T(c["inputs"]), emb, T(c["offsets"]), out, B, D, C, L, S, 16, dy, 0, False, 0)
            rng = np.random.default_rng(cs["seed"] + 100)
            g = (rng.normal(size=(L, B, C)) * 0.01).astype(np.float32)
            gemb = torch.zeros_like(emb)
            gin = torch.zeros(B, D, device=dev, dtype=dt)
            ge.grid_encode_backward(T(g).to(dt), T(c["inputs"]), emb, T(c["offsets"]), gemb, B, D, C, L, S, 16, dy, gin, 0,
                                    False, 0)
            nz = gemb.float().abs().sum(1).nonzero()[:, 0]
            # the device's exp2f for the per-level scale (see lnb_oracle.c: CUDA exp2f != glibc exp2f by 1 ulp)
            lv = torch.arange(L, device=dev, dtype=torch.float32) * torch.tensor(S, device=dev, dtype=torch.float32)
            level_scales = torch.exp2(lv) * 16.0 - 1.0
            save(f"ref_cuda_grid{i}.npz", cfg=np.array(repr(cs)), out=out.float(), dy_dx=dy.float(), grad=g,
                 level_scales=level_scales,
                 grad_rows=nz.int(), grad_vals=gemb.float()[nz], grad_inputs=gin.float())

    fe = refcuda.load("_freqencoder")
    if fe is not None:
        rng = np.random.default_rng(41)
        x = rng.uniform(-1, 1, size=(300, 3)).astype(np.float32)
        out = torch.empty(300, 75, device=dev)
        fe.freq_encode_forward(T(x), 300, 3, 12, 75, out)
        g = rng.normal(size=(300, 75)).astype(np.float32)
        gi = torch.zeros(300, 3, device=dev)
        fe.freq_encode_backward(T(g), out, 300, 3, 12, 75, gi)
        save("ref_cuda_freq.npz", x=x, out=out, g=g, grad_inputs=gi)

    sh = refcuda.load("_shencoder")
    if sh is not None:
        rng = np.random.default_rng(42)
        x = rng.uniform(-1, 1, size=(200, 3)).astype(np.float32)
        x[:100] = cases.unit(x[:100])
        res = {}
        for deg in (1, 2, 4, 6, 8):
            out = torch.empty(200, deg * deg, device=dev)
            dy = torch.empty(200, 3 * deg * deg, device=dev)
            sh.sh_encode_forward(T(x), out, 200, 3, deg, dy)
            res[f"out{deg}"], res[f"dy{deg}"] = out, dy
        save("ref_cuda_sh.npz", x=x, **res)

    ff = refcuda.load("_ffmlp")
    if ff is not None:
        for i, (B, ind, nl) in enumerate([(256, 32, 2), (384, 96, 2), (128, 64, 3)]):
            c = cases.ffmlp_case(50 + i, B, ind, 64, nl, 16)
            ff.allocate_splitk(nl + 1)
            x, w, g = T(c["x"]), T(c["w"]), T(c["g"])
            fb = torch.empty(nl, B, 64, device=dev, dtype=torch.half)
            out = torch.empty(B, 16, device=dev, dtype=torch.half)
            ff.ffmlp_forward(x, w, B, ind, 16, 64, nl, 0, 6, fb, out)
            bb = torch.zeros(nl, B, 64, device=dev, dtype=torch.half)
            gi = torch.zeros(B, ind, device=dev, dtype=torch.half)
            gw = torch.zeros_like(w)
            ff.ffmlp_backward(g, x, w, fb, B, ind, 16, 64, nl, 0, 6, True, bb, gi, gw)
            torch.cuda.synchronize()
            save(f"ref_cuda_ffmlp{i}.npz", cfg=np.array(repr((B, ind, nl))), out=out.float(), fb=fb.float(),
                 grad_inputs=gi.float(), grad_weights=gw.float(), bb=bb.float())
    print("reference modules used:", refcuda.available())


if __name__ == "__main__":
    main()
