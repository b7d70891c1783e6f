"""How many of the marched samples does a ray need?  After `warm` training steps of the bench workload: per ray the number
of marched samples and the index at which compositing stops (T < T_thresh); prints the distribution and the rows a chunked
forward (evaluate a ray's samples in rounds of R_0, R_1, ... and stop as soon as the ray is opaque) would evaluate."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence

dev = torch.device("cuda:0")
for warm in (16, 300, 1200):
    cfg = FieldConfig()
    seq = SyntheticLidarSequence(n_frames=8, device=dev)
    eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 256)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    for it in range(warm):
        eng.set_batch(*seq.sample_batch(4096, frame=it % 8, generator=gen, device=dev))
        eng.train_step(use_graph=False)
    eng.flush()
    eng.set_batch(*seq.sample_batch(4096, frame=0, generator=gen, device=dev))
    eng._forward_backward()
    torch.cuda.synchronize()
    rays = eng.rays.cpu().numpy()
    sig = eng.sigma.cpu().numpy()
    dl = eng.deltas.cpu().numpy()[:, 0]
    cnt, stop = [], []
    for idx, off, c in rays:
        if c == 0:
            cnt.append(0); stop.append(0); continue
        a = 1 - np.exp(-sig[off:off + c] * dl[off:off + c])
        T = np.cumprod(1 - a)
        k = np.nonzero(T < cfg.T_thresh)[0]
        cnt.append(c); stop.append(int(k[0]) + 1 if len(k) else c)
    cnt, stop = np.array(cnt), np.array(stop)
    print(f"after {warm} steps: marched {cnt.sum()} ({cnt.mean():.1f}/ray), needed {stop.sum()} ({stop.mean():.1f}/ray), "
          f"rays that never stop: {(stop == cnt).mean() * 100:.1f} %")
    print("   stop index percentiles 10/25/50/75/90/99:", np.percentile(stop, [10, 25, 50, 75, 90, 99]).astype(int).tolist(),
          " marched 50/90/99:", np.percentile(cnt, [50, 90, 99]).astype(int).tolist())
    for sched in ([32, 64, 128], [32, 64], [64, 128], [32, 96], [16, 32, 64, 128], [64]):
        bounds = [0] + sched + [10 ** 9]
        rows = 0
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            alive = stop > lo                       # the ray still needs samples beyond lo
            rows += np.minimum(cnt[alive], hi).clip(min=lo).sum() - lo * alive.sum()
        print(f"   rounds at {sched}: {rows} rows evaluated ({rows / cnt.sum() * 100:.0f} % of marched, {len(sched) + 1} rounds)")
