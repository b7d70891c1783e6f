"""Per-kernel device time INSIDE the captured CUDA graph of the training step (event-record nodes between the
kernels), next to the eager numbers bench.py reports - do kernels run slower back to back than with launch gaps?

    [LNB_COMPACT_BACKWARD=0] python scripts/diag_graph_times.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf import engine as E
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence

dev = torch.device("cuda:0")
cfg = FieldConfig()
if os.environ.get("LNB_COMPACT_BACKWARD") == "0":
    cfg.compact_backward = False
if os.environ.get("LNB_LATE_GRAD_ZERO") == "0":
    cfg.late_grad_zero = False
seq = SyntheticLidarSequence(n_frames=8, device=dev)
eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 64)
eng.seed_occupancy_from_points(seq.surface_points())
gen = torch.Generator().manual_seed(0)
batches = []
for i in range(16):
    batches.append(seq.sample_batch(4096, generator=gen, device=dev))
for it in range(260):
    eng.set_batch(*batches[it % 16])
    eng.train_step(use_graph=it > 8)
    if it == 5:
        eng.fit_sample_budget(1.6)
torch.cuda.synchronize()

marks = []


class Proxy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, n):
        fn = getattr(self._real, n)
        if not n.startswith("lnb_") or n in ("lnb_field_supported",):
            return fn

        def timed(*a):
            ext = torch.cuda.is_current_stream_capturing()
            e0 = torch.cuda.Event(enable_timing=True, external=ext)
            e1 = torch.cuda.Event(enable_timing=True, external=ext)
            e0.record()
            r = fn(*a)
            e1.record()
            marks.append((n, e0, e1))
            return r
        return timed


real = E.lib
E.lib = Proxy(real)
try:
    eng._graph = None
    t0 = torch.cuda.Event(enable_timing=True, external=True)
    t1 = torch.cuda.Event(enable_timing=True, external=True)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng._forward_backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    marks.clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        t0.record()
        eng._forward_backward()
        t1.record()
finally:
    E.lib = real
acc = {}
tot = 0.0
iters = 30
for it in range(iters + 3):
    eng.set_batch(*batches[it % 16])
    g.replay()
    eng._optimizer()
    torch.cuda.synchronize()
    if it >= 3:
        tot += t0.elapsed_time(t1) * 1e3
        for n, e0, e1 in marks:
            acc[n] = acc.get(n, 0.0) + e0.elapsed_time(e1) * 1e3
print(f"compact_backward={cfg.compact_backward} late_grad_zero={cfg.late_grad_zero} samples={eng.samples_last_step()} live={int(eng.counter[2])}")
print(f"graph forward+backward: {tot / iters:.1f} us")
s = 0
for n, v in acc.items():
    print(f"  {n:40s} {v / iters:7.1f} us")
    s += v / iters
print(f"  sum {s:.1f} us")

# the same kernels launched eagerly (Python launch gaps between them), ordinary events
marks.clear()
E.lib = Proxy(real)
acc2 = {}
try:
    for it in range(iters + 3):
        eng.set_batch(*batches[it % 16])
        marks.clear()
        eng._forward_backward()
        eng._optimizer()
        torch.cuda.synchronize()
        if it >= 3:
            for n, e0, e1 in marks:
                acc2[n] = acc2.get(n, 0.0) + e0.elapsed_time(e1) * 1e3
finally:
    E.lib = real
print("eager, same state:")
for n, v in acc2.items():
    print(f"  {n:40s} {v / iters:7.1f} us")
