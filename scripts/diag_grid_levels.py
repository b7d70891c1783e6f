"""Cost of the hash-grid scatter per level window, in two training states (LiDAR prior only vs network-refreshed
occupancy).  python scripts/diag_grid_levels.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf import engine as E
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence

dev = torch.device("cuda:0")


class Recorder:
    def __init__(self, real):
        self._real, self.calls = real, {}

    def __getattr__(self, n):
        fn = getattr(self._real, n)
        if not n.startswith("lnb_grid_encode_backward"):
            return fn

        def rec(*a):
            self.calls[n] = (fn, a)
            return fn(*a)
        return rec


def run(refresh):
    cfg = FieldConfig(grid_update_interval=16 if refresh else 0)
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 160)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    for it in range(260):
        eng.set_batch(*seq.sample_batch(4096, frame=it % 8, generator=gen, device=dev))
        eng.train_step(use_graph=False)
    real = E.lib
    rec = Recorder(real)
    E.lib = rec
    try:
        eng._forward_backward()
    finally:
        E.lib = real
    torch.cuda.synchronize()
    (name, (fn, a)), = rec.calls.items()
    n, nl = int(eng.counter[0]), int(eng.counter[2])
    print(f"state: refresh={refresh} samples={n} live={nl} kernel={name}")
    for lo, hi in ((0, 16), (0, 1), (1, 2), (2, 4), (4, 6), (6, 9), (9, 12), (12, 16)):
        os.environ["LNB_GRID_BWD_LEVELS"] = f"{lo},{hi}"
        best = 1e9
        for it in range(8):
            eng.g_table.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            assert fn(*a) == 0
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        print(f"  levels [{lo:2d},{hi:2d}): {best:7.1f} us")
    os.environ.pop("LNB_GRID_BWD_LEVELS")


run(False)
run(True)
