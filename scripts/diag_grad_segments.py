"""Norms of the gradient segments (hash table | density MLP | head) after ONE full-size step, for the backward variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
dev = torch.device("cuda:0")
seq = SyntheticLidarSequence(n_frames=8, device=dev)
for n_rays in (256, 4096):
    for over in (dict(), dict(compact_backward=False), dict(fused_composite=False), dict(fused_field=False), dict(fused_gather=False)):
        cfg = FieldConfig(grid_update_interval=0, perturb=False, **over)
        eng = LidarFieldEngine(cfg, n_rays, device=dev, sample_budget=n_rays * 256)
        eng.seed_occupancy_from_points(seq.surface_points())
        gen = torch.Generator().manual_seed(0)
        eng.set_batch(*seq.sample_batch(n_rays, frame=0, generator=gen, device=dev))
        eng.G.zero_()
        eng._forward_backward()
        torch.cuda.synchronize()
        a, b, n = eng.n_table, eng.n_table + eng.n_sigma, eng.n_params
        G = eng.G
        hw = G[b:n]
        nin = 64 * cfg.head_in_dim
        print(f"rays {n_rays} {over}: samples {int(eng.counter[0])} live {int(eng.counter[2])} |g_table| {float(G[:a].norm()):.4e} "
              f"|g_sigma_w| {float(G[a:b].norm()):.4e} |g_head_w| {float(hw.norm()):.4e} "
              f"(W_in {float(hw[:nin].norm()):.3e}, W_hid {float(hw[nin:nin + 4096].norm()):.3e}, W_out {float(hw[nin + 4096:].norm()):.3e}) loss {float(eng.loss_acc):.4f}")
