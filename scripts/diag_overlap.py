"""Do two kernels of the training step overlap usefully when they share the GPU?

    python scripts/diag_overlap.py

Records the C-ABI calls of one eager step of the bench workload, then replays chosen kernels alone and pairwise on two
streams (CUDA events on a third, joined stream) and prints t(A), t(B), t(A || B).  The pairs are the candidates for
chunk pipelining / cross-step overlap (DESIGN.md "what comes next"); data hazards are irrelevant for timing because
every kernel rewrites the values that are already in its output buffers.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch   # noqa: E402
from lidar_nerf_b200.nerf import engine as E   # noqa: E402
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig   # noqa: E402
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence       # noqa: E402
from lidar_nerf_b200 import backend   # noqa: E402


class Recorder:
    def __init__(self, real):
        self._real, self.calls = real, []

    def __getattr__(self, n):
        fn = getattr(self._real, n)
        if not n.startswith("lnb_") or n in ("lnb_field_supported", "lnb_strerror"):
            return fn

        def rec(*a):
            self.calls.append((n, fn, a))
            return fn(*a)
        return rec


def main():
    dev = torch.device("cuda:0")
    cfg = FieldConfig(grid_update_interval=0)
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 64)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    for i in range(4):
        ro, rd, gt = seq.sample_batch(4096, generator=gen, device=dev)
        eng.set_batch(ro, rd, gt)
        eng.train_step(use_graph=False)
        eng.fit_sample_budget()
    eng.update_density_grid(full=True)
    eng.train_step(use_graph=False)
    eng.fit_sample_budget(1.6)
    torch.cuda.synchronize()
    real = E.lib
    rec = Recorder(real)
    E.lib = rec
    try:
        eng._forward_backward()
    finally:
        E.lib = real
    torch.cuda.synchronize()
    print("samples:", eng.samples_last_step(), "M:", eng.M)
    calls = {}
    for n, fn, a in rec.calls:
        key = n if n not in calls else n + "#2"
        calls[key] = (fn, a)
    print("recorded:", list(calls))

    sa, sb = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def launcher(key):
        if key == "adam":
            def go(stream):
                with torch.cuda.stream(stream):
                    backend.adam_step(eng.P, eng.G, eng.m, eng.v, eng.Ph, 0.0, 0.9, 0.99, 1e-15, 10, grad_scale=1.0, zero_grad=False)
            return go
        fn, a = calls[key]

        def go(stream):
            rc = fn(*a[:-1], C.c_void_p(stream.cuda_stream))
            assert rc == 0, (key, rc)
        return go

    def timed(keys, iters=20):
        gos = [launcher(k) for k in keys]
        streams = [sa, sb][:len(keys)]
        main_s = torch.cuda.current_stream()
        best = 1e9
        for it in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main_s)
            for s in streams:
                s.wait_stream(main_s)
            for g, s in zip(gos, streams):
                g(s)
            for s in streams:
                main_s.wait_stream(s)
            e1.record(main_s)
            torch.cuda.synchronize()
            if it >= 3:
                best = min(best, e0.elapsed_time(e1) * 1e3)
        return best

    names = {"march": "lnb_march_rays_train_ex", "grid_fwd": "lnb_grid_encode_forward_ex", "field_fwd": "lnb_field_forward",
             "comp": "lnb_lidar_composite_step", "head_bwd": "lnb_field_head_backward",
             "sigma_bwd": "lnb_ffmlp_backward_accumulate", "grid_bwd": "lnb_grid_encode_backward_ex", "adam": "adam",
             "ray_terms": "lnb_field_ray_terms"}
    solo = {k: timed([v]) for k, v in names.items() if v in calls or v == "adam"}
    for k, v in solo.items():
        print(f"solo {k:10s} {v:7.1f} us")
    pairs = [("grid_fwd", "field_fwd"), ("sigma_bwd", "grid_bwd"), ("head_bwd", "sigma_bwd"), ("head_bwd", "grid_bwd"),
             ("adam", "march"), ("adam", "grid_fwd"), ("adam", "field_fwd"), ("adam", "grid_bwd"), ("march", "grid_bwd"),
             ("grid_fwd", "grid_bwd"), ("field_fwd", "head_bwd")]
    for a, b in pairs:
        if a in solo and b in solo:
            t = timed([names[a], names[b]])
            t2 = timed([names[b], names[a]])
            print(f"pair {a:10s} || {b:10s}: {t:7.1f} us (launch order swapped {t2:7.1f}) vs sum {solo[a] + solo[b]:7.1f}, max {max(solo[a], solo[b]):7.1f}")


if __name__ == "__main__":
    main()
