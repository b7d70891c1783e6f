"""Where does the persistent gather + MLP forward kernel spend its time?  Times lnb_field_fused_forward alone (CUDA
events, steady-state sample set of the bench workload) against the two-kernel path, and with parts of it switched off
(LNB_FUSED_DBG_LIVE bits: 1 = gather without table loads, 2 = no enc / saved-activation stores, 4 = no epilogue math).
    python scripts/diag_fused_fwd.py [steps-of-training-before-measuring]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf import engine as E
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence

dev = torch.device("cuda:0")
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 300


class Recorder:
    def __init__(self, real, names):
        self._real, self.calls, self.names = real, {}, names

    def __getattr__(self, n):
        fn = getattr(self._real, n)
        if n not in self.names:
            return fn

        def rec(*a):
            self.calls[n] = (fn, a)
            return fn(*a)
        return rec


def best_of(fn, a, reps=10):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert fn(*a) == 0
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


def main():
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    calls = {}
    for fg in (True, False):
        cfg = FieldConfig(fused_gather=fg)
        eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 256)
        eng.seed_occupancy_from_points(seq.surface_points())
        gen = torch.Generator().manual_seed(0)
        for it in range(warm):
            eng.set_batch(*seq.sample_batch(4096, frame=it % 8, generator=gen, device=dev))
            eng.train_step(use_graph=False)
        real = E.lib
        rec = Recorder(real, ("lnb_field_fused_forward", "lnb_grid_encode_forward_ex", "lnb_field_forward"))
        E.lib = rec
        try:
            eng._forward_backward()
        finally:
            E.lib = real
        torch.cuda.synchronize()
        n = int(eng.counter[0])
        print(f"fused_gather={fg}: samples={n} ({n / 4096:.1f}/ray), tiles={(n + 127) // 128}")
        for name, (fn, a) in rec.calls.items():
            if name == "lnb_field_fused_forward":
                for dbg, what in ((0, "full"), (1, "gather without table loads"), (2, "no enc/activation stores"),
                                  (3, "no loads, no stores"), (4, "no epilogue math/stores"), (7, "skeleton: barriers + MMAs + gather index math"),
                                  (15, "protocol only: barriers + MMAs"), (9, "MLP only: no gather work at all"),
                                  (6, "gather only: no epilogue math/stores")):
                    os.environ["LNB_FUSED_DBG_LIVE"] = str(dbg)
                    print(f"  {name} [{what}]: {best_of(fn, a):7.1f} us")
                os.environ["LNB_FUSED_DBG_LIVE"] = "0"
            else:
                print(f"  {name}: {best_of(fn, a):7.1f} us")
        calls[fg] = eng
    return 0


if __name__ == "__main__":
    sys.exit(main())
