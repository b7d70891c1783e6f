"""Where does the data-parallel step spend the time it adds over the single-GPU step?  Run under torchrun (2+ ranks):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/diag_dp_exchange.py
CUDA events around the two graph replays of a step (main stream) and around barrier / exchange kernel / barrier (the
communication stream), averaged over 60 steady-state steps, printed per rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench as B
from lidar_nerf_b200.nerf import engine as E
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    wl = B.WORKLOADS[2]
    N = wl["rays"]
    cfg = FieldConfig(**wl["field"])
    if os.environ.get("LNB_MULTICAST_EXCHANGE") == "1":
        cfg.multicast_exchange = True
    seq = SyntheticLidarSequence(n_frames=8, device=dev, **wl["seq"])
    eng = LidarFieldEngine(cfg, N, device=dev, sample_budget=N * 258)
    eng.seed_occupancy_from_points(seq.surface_points())
    pool = B.make_pool(seq, N, 64, seed=1000 + rank, device=dev)
    for i in range(3):
        eng.set_batch_packed(pool[i])
        eng.train_step(use_graph=False)
    for i in range(480):
        eng.set_batch_packed(pool[i % 64])
        eng.train_step()
    eng.flush()
    torch.cuda.synchronize()
    dist.barrier()
    cfg.grid_update_interval = 0
    ev = []

    def mark(tag):
        e = torch.cuda.Event(enable_timing=True)
        e.record()                       # on the CURRENT stream (main or communication)
        ev.append((tag, e))

    class BarrierProxy:
        def __init__(self, real):
            self._r = real

        def __getattr__(self, n):
            return getattr(self._r, n)

        def barrier(self, channel=0, timeout_ms=0):
            mark(f"bar{channel}_in")
            self._r.barrier(channel=channel, timeout_ms=timeout_ms)
            mark(f"bar{channel}_out")

    assert eng._peer is not None, "peer-memory exchange not set up"
    eng._peer["hG"] = BarrierProxy(eng._peer["hG"])
    ga, gb = eng._graph, eng._graph_b

    class GraphProxy:
        def __init__(self, real, tag):
            self._r, self._t = real, tag

        def replay(self):
            mark(self._t + "_in")
            self._r.replay()
            mark(self._t + "_out")

    eng._graph, eng._graph_b = GraphProxy(ga, "A"), GraphProxy(gb, "B")
    steps = 60
    for i in range(steps):
        eng.set_batch_packed(pool[i % 64])
        mark("step")
        eng.train_step()
    eng.flush()
    mark("step")
    torch.cuda.synchronize()
    # per step: offsets relative to the step's first event
    starts = [i for i, (t, _) in enumerate(ev) if t == "step"]
    acc = {}
    for a, b in zip(starts[5:-1], starts[6:]):          # skip the first steps
        t0 = ev[a][1]
        for tag, e in ev[a + 1:b]:
            acc.setdefault(tag, []).append(t0.elapsed_time(e) * 1e3)
        acc.setdefault("next_step", []).append(t0.elapsed_time(ev[b][1]) * 1e3)
    order = ["A_in", "A_out", "B_in", "B_out", "bar0_in", "bar0_out", "bar1_in", "bar1_out", "next_step"]
    line = f"rank {rank}/{world} (us after the step's first event; mean over {len(acc['next_step'])} steps): " + \
        ", ".join(f"{k}={sum(acc[k]) / len(acc[k]):.0f}" for k in order if k in acc)
    d = {k: sum(v) / len(v) for k, v in acc.items()}
    line += (f"\n   march graph {d['A_out'] - d['A_in']:.0f} us | wait for the previous exchange {d['B_in'] - d['A_out']:.0f} | "
             f"forward+backward graph {d['B_out'] - d['B_in']:.0f} | barrier {d['bar0_out'] - d['bar0_in']:.0f} | "
             f"exchange kernel {d['bar1_in'] - d['bar0_out']:.0f} | barrier {d['bar1_out'] - d['bar1_in']:.0f} | step {d['next_step']:.0f}")
    for r in range(world):
        if r == rank:
            print(line, flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
