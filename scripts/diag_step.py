"""GPU diagnostic: where does the step time go under different launch/sync regimes? (writes gpurun_out/diag.txt)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
from lidar_nerf_b200._lib import lib, u32, f32, i32, vp
import bench

dev = torch.device("cuda:0")
N = 4096
cfg = FieldConfig()
seq = SyntheticLidarSequence(n_frames=8, device=dev)
eng = LidarFieldEngine(cfg, N, device=dev, sample_budget=N * 64)
eng.seed_occupancy_from_points(seq.surface_points())
pool = bench.make_pool(seq, N, 32, 1000, dev)


def load(b):
    eng.rays_o.copy_(b[:, 0:3]); eng.rays_d.copy_(b[:, 3:6]); eng.gt.copy_(b[:, 6:9])


cfg.grid_update_interval = 0
for i in range(3):
    load(pool[i]); eng.train_step(use_graph=False); eng.fit_sample_budget()
eng.update_density_grid(full=True)
load(pool[0]); eng.train_step(use_graph=False); eng.fit_sample_budget(1.6)
print("M", eng.M, "produced", eng.samples_last_step())


def timeit(fn, steps=48, sync_each=False):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
        if sync_each:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


def step_graph(i):
    load(pool[i % 32]); eng.train_step(True)


def step_eager(i):
    load(pool[i % 32]); eng.train_step(False)


for _ in range(4):
    step_graph(0)
out = []
out.append(("graph, no sync", timeit(step_graph)))
out.append(("graph, sync each step", timeit(step_graph, sync_each=True)))
out.append(("eager, no sync", timeit(step_eager)))
out.append(("eager, sync each step", timeit(step_eager, sync_each=True)))
out.append(("graph only (no adam, no load)", timeit(lambda i: eng._graph.replay())))
out.append(("adam only", timeit(lambda i: eng._optimizer())))
out.append(("graph only again", timeit(lambda i: eng._graph.replay())))
# isolated grid backward, repeated (G warm in L2 after the first)
p = lambda t: vp(t.data_ptr())
s = vp(torch.cuda.current_stream().cuda_stream)
c = cfg


def gbwd(i):
    lib.lnb_grid_encode_backward_ex(p(eng.g_enc), p(eng.xyzs), p(eng.table_h), p(eng.offsets), p(eng.g_table), u32(eng.M),
                                    u32(3), u32(2), u32(16), f32(eng.S), u32(16), vp(0), vp(0), u32(0), i32(0), u32(0),
                                    i32(1), i32(1), f32(c.bound), i32(1), p(eng.counter), s)


out.append(("grid_bwd back-to-back (warm)", timeit(gbwd, 20)))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def gbwd_cold(i):
    flush.zero_()
    gbwd(i)


t_flush = timeit(lambda i: flush.zero_(), 20)
out.append(("grid_bwd after L2 flush (minus flush)", timeit(gbwd_cold, 20) - t_flush))
def gbwd_after_adam(i):
    eng._optimizer(); gbwd(i)
t_adam = timeit(lambda i: eng._optimizer(), 20)
out.append(("grid_bwd right after adam (minus adam)", timeit(gbwd_after_adam, 20) - t_adam))
os.environ_copy = None
for k, v in out:
    print(f"{k:45s} {v:10.1f} us")
with open("gpurun_out/diag.txt", "w") as f:
    for k, v in out:
        f.write(f"{k:45s} {v:10.1f} us\n")
