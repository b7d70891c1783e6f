"""In-kernel timeline of the persistent gather + MLP forward kernel (diagnostic build: `python lidar-nerf_b200/build.py --trace`).

    LNB200_LIB=lidar-nerf_b200/lib/liblnb200_trace.so python scripts/diag_fwd_trace.py [dbg-bits ...]

For CTA 0: SM-clock timeline of one thread per role (gather warp 0, epilogue group 0 / 1, MMA warp) over two steady-state
tiles, for each LNB_FUSED_DBG_LIVE setting given (default 0 = the full kernel, 15 = barriers + MMAs only)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LNB200_LIB", os.path.join(ROOT, "lidar-nerf_b200", "lib", "liblnb200_trace.so"))
import torch         # noqa: E402
from lidar_nerf_b200 import _lib   # noqa: E402
from lidar_nerf_b200.nerf import engine as E   # noqa: E402
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig   # noqa: E402
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence       # noqa: E402

ROLE = {0: "epi0", 1: "epi1", 2: "mma", 3: "gather"}
EPI = {1: "tile: wait xfull", 2: "xfull seen", 3: "wait done", 4: "done seen", 5: "math done", 6: "published", 7: "copied out"}
GAT = {1: "tile top", 2: "past named bar", 3: "stage free", 4: "levels done", 5: "arrived xfull"}


def read(raw):
    fn = raw.lnb_debug_fwd_trace_fused
    fn.restype = C.c_int
    buf = (C.c_ulonglong * (4 * 4096))()
    assert fn(buf, C.c_int(1)) == 0
    ev = []
    for role in range(4):
        for i in range(4096):
            v = buf[role * 4096 + i]
            if v == 0:
                break
            tag, clk = v >> 44, v & ((1 << 44) - 1)
            ev.append((clk, role, (tag >> 8) & 255, tag & 255))
    ev.sort()
    return ev


def show(ev, first_tile=8, n_tiles=2):
    # window: from the gather's "tile top" of tile `first_tile` to that of tile first_tile + n_tiles
    tops = {arg: clk for clk, role, e, arg in ev if role == 3 and e == 1}
    if first_tile + n_tiles not in tops:
        first_tile = max(0, max(tops) - n_tiles - 1)
    lo, hi = tops[first_tile], tops[first_tile + n_tiles]
    print(f"  gather tiles {first_tile}..{first_tile + n_tiles - 1} of CTA 0: {hi - lo} cycles ({(hi - lo) / n_tiles:.0f} per tile); "
          f"whole CTA: {ev[-1][0] - ev[0][0]} cycles for {max(tops) + 1} tiles")
    last = {}
    for clk, role, e, arg in ev:
        if role == 2:
            q, issued = (e - 1) // 2, (e - 1) % 2
            what = f"slot{q} step{arg} " + ("issued" if issued else "ready seen")
        elif role == 3:
            what = f"{GAT.get(e, e)} k={arg}"
        else:
            what = f"{EPI.get(e, e)} " + (f"k={arg}" if e <= 2 else f"ph{arg}")
        dt = clk - last.get(role, clk)
        last[role] = clk
        if lo <= clk < hi:
            print(f"{clk - lo:8d} " + " " * (30 * role) + f"[{ROLE[role]}] {what} (+{dt})")


def main():
    dev = torch.device("cuda:0")
    modes = [int(x) for x in sys.argv[1:]] or [0, 15]
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    eng = LidarFieldEngine(FieldConfig(), 4096, device=dev, sample_budget=4096 * 256)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    for it in range(300):
        eng.set_batch(*seq.sample_batch(4096, frame=it % 8, generator=gen, device=dev))
        eng.train_step(use_graph=False)
    torch.cuda.synchronize()
    raw = C.CDLL(os.environ["LNB200_LIB"])
    for dbg in modes:
        os.environ["LNB_FUSED_DBG_LIVE"] = str(dbg)
        read(raw)                       # reset
        eng._forward_backward()
        torch.cuda.synchronize()
        print(f"==== LNB_FUSED_DBG_LIVE={dbg} ====")
        show(read(raw))
    os.environ["LNB_FUSED_DBG_LIVE"] = "0"
    return 0


if __name__ == "__main__":
    sys.exit(main())
