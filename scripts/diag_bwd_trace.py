"""In-kernel timeline of the MLP backward (diagnostic build: `python lidar-nerf_b200/build.py --trace`).

    LNB200_LIB=lidar-nerf_b200/lib/liblnb200_trace.so python scripts/diag_bwd_trace.py

Prints, for CTA 0 of the LiDAR-head backward and of the density-MLP backward, the SM-clock timeline of one thread per
role (compute warpgroup 0/1, MMA warp) over the first tiles of a bench-sized step.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LNB200_LIB", os.path.join(ROOT, "lidar-nerf_b200", "lib", "liblnb200_trace.so"))
import numpy as np   # noqa: E402
import torch         # noqa: E402
from lidar_nerf_b200 import _lib   # noqa: E402
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig   # noqa: E402
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence       # noqa: E402

EV = {0: "top", 1: "landed", 2: "arrive0", 3: "done", 4: "tmem", 5: "sts", 6: "arrive", 7: "tile_end", 8: "synced"}
MMA = {1: "ready g0", 2: "ready g1", 3: "issued g0", 4: "issued g1"}


def read(name):
    fn = getattr(_lib.lib, f"lnb_debug_bwd_trace_{name}")
    buf = (C.c_ulonglong * 16384)()
    n = fn(buf, C.c_uint32(16384), C.c_int(1))
    ev = []
    for i in range(n):
        v = buf[i]
        tag, clk = v >> 44, v & ((1 << 44) - 1)
        ev.append((clk, tag >> 12, (tag >> 4) & 0xff, tag & 15))
    ev.sort()
    return ev


def show(name, ev, limit=140):
    if not ev:
        print(name, "no events")
        return
    t0 = ev[0][0]
    print(f"==== {name}: {len(ev)} events; cycles relative to the first event ====")
    last = {}
    for clk, role, e, arg in ev[:limit]:
        what = (MMA.get(e, str(e)) + f" ph{arg}") if role == 2 else (EV.get(e, str(e)) + (f" L{arg}" if e in (3, 4, 5, 6) else ""))
        dt = clk - last.get(role, clk)
        last[role] = clk
        col = {0: 0, 1: 34, 2: 68}[role]
        print(f"{clk - t0:9d} " + " " * col + f"[{'wg0' if role == 0 else 'wg1' if role == 1 else 'mma'}] {what} (+{dt})")


def main():
    dev = torch.device("cuda:0")
    cfg = FieldConfig()
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 64)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    cfg.grid_update_interval = 0
    for i in range(4):
        ro, rd, gt = seq.sample_batch(4096, generator=gen, device=dev)
        eng.set_batch(ro, rd, gt)
        eng.train_step(use_graph=False)
        eng.fit_sample_budget()
    torch.cuda.synchronize()
    read("head"), read("generic")          # reset
    eng.G.zero_()
    eng._forward_backward()
    torch.cuda.synchronize()
    print("samples this step:", eng.samples_last_step())
    show("LiDAR-head backward (CTA 0)", read("head"))
    show("density-MLP backward (CTA 0)", read("generic"))


if __name__ == "__main__":
    main()
