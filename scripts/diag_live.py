"""How many marched samples still carry a gradient after compositing (early stop at T < 1e-4), as training proceeds?
Decides whether compacting the backward pass to live samples pays.  python scripts/diag_live.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence

dev = torch.device("cuda:0")
cfg = FieldConfig()
seq = SyntheticLidarSequence(n_frames=8, device=dev)
eng = LidarFieldEngine(cfg, 4096, device=dev, sample_budget=4096 * 200)
eng.seed_occupancy_from_points(seq.surface_points())
gen = torch.Generator().manual_seed(0)
for it in range(3001):
    ro, rd, gt = seq.sample_batch(4096, generator=gen, device=dev)
    eng.set_batch(ro, rd, gt)
    eng.train_step(use_graph=False)
    if it in (0, 10, 50, 100, 200, 400, 800, 1500, 3000):
        n = int(eng.counter[0].item())
        live = (eng.g_sigma[:n] != 0) | (eng.g_rgb[:n] != 0).any(-1)
        genc = (eng.g_enc[:n] != 0).any(-1)
        rays = eng.rays.cpu()
        print(f"step {it:5d}: samples {n:7d} ({n / 4096:.1f}/ray)  live after composite {float(live.float().mean()):.3f}  "
              f"nonzero g_enc rows {float(genc.float().mean()):.3f}  loss {eng.read_loss():.4f}", flush=True)
