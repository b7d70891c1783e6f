"""Does the LiDAR head learn?  Trains the bench workload (a) with the fused engine and (b) through the module API
(NeRFNetwork.render fused Function -> torch loss -> GradScaler -> torch Adam) and prints the three loss components
(depth L1 [m], ray-drop MSE, intensity MSE) on the training batches every 100 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig
from lidar_nerf_b200.nerf.network_tcnn import NeRFNetwork
from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence
from lidar_nerf_b200 import raymarching as rmw

dev = torch.device("cuda:0")
N, STEPS = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seq = SyntheticLidarSequence(n_frames=8, device=dev)


def comps(depth_abs, image, gt):
    m = gt[:, 0]
    return (float(((depth_abs - gt[:, 2]).abs() * m).sum() / m.sum()) / seq.scale,
            float(((image[:, 0] - m) ** 2).mean()), float((((image[:, 1] - gt[:, 1]) ** 2) * m).sum() / m.sum()))


def run_engine(**over):
    cfg = FieldConfig(**over)
    eng = LidarFieldEngine(cfg, N, device=dev, sample_budget=N * 256)
    eng.seed_occupancy_from_points(seq.surface_points())
    gen = torch.Generator().manual_seed(0)
    acc = [0.0, 0.0, 0.0]
    hs = slice(eng.n_table + eng.n_sigma, eng.n_params)
    ss = slice(eng.n_table, eng.n_table + eng.n_sigma)
    w_prev, s_prev = eng.P[hs].clone(), eng.P[ss].clone()
    for it in range(1, STEPS + 1):
        ro, rd, gt = seq.sample_batch(N, frame=it % 8, generator=gen, device=dev)
        eng.set_batch(ro, rd, gt)
        eng.train_step(use_graph=True)
        if it % 100 == 0:
            g = eng.G[hs] / cfg.loss_scale
            m, v = eng.m[hs], eng.v[hs]
            upd = (m / (v.sqrt() + 1e-15)).abs()
            print(f"     head: |dw| over 100 steps {float((eng.P[hs] - w_prev).norm()):.4f}  |g| {float(g.norm()):.3e} max|g| {float(g.abs().max()):.3e} "
                  f"mean|m/sqrt(v)| {float(upd.mean()):.3e}  nan {int(torch.isnan(eng.P[hs]).sum())}   sigma-net |dw| {float((eng.P[ss] - s_prev).norm()):.4f}")
            w_prev, s_prev = eng.P[hs].clone(), eng.P[ss].clone()
        c = comps(torch.addcmul(eng.depth, eng.t0, eng.ws), eng.image, gt)
        acc = [a + b for a, b in zip(acc, c)]
        if it % 100 == 0:
            print(f"  engine{over} step {it:5d}: depth L1 {acc[0] / 100:.3f} m  raydrop MSE {acc[1] / 100:.4f}  intensity MSE {acc[2] / 100:.4f}"
                  f"  |w_head| {float(eng.P[eng.n_table + eng.n_sigma:eng.n_params].norm()):.3f}", flush=True)
            acc = [0.0, 0.0, 0.0]


def run_b2(mode="fused", **rkw):
    torch.manual_seed(0)
    net = NeRFNetwork(encoding="hashgrid", desired_resolution=32768, log2_hashmap_size=19, n_features_per_level=2, num_layers=2,
                      hidden_dim=64, geo_feat_dim=15, bound=1, density_scale=1, min_near=seq.scale, min_near_lidar=seq.scale,
                      density_thresh=10, bg_radius=-1).to(dev)
    net.train()
    H = net.grid_size
    pts = seq.surface_points()
    offs = torch.stack(torch.meshgrid(*([torch.arange(-1, 2, device=dev)] * 3), indexing="ij"), -1).reshape(-1, 3)
    cell = torch.clamp((0.5 * (pts + 1) * H).long(), 0, H - 1)
    cell = torch.unique((cell[:, None, :] + offs[None]).reshape(-1, 3).clamp(0, H - 1), dim=0)
    prior = torch.zeros(1, H ** 3, device=dev)
    prior[0, rmw.morton3D(cell.int()).long()] = 1.0
    rmw.packbits(prior, 0.5, net.density_bitfield)
    net.grid_update_interval = 0
    opt = torch.optim.Adam(net.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    gen = torch.Generator().manual_seed(0)
    acc = [0.0, 0.0, 0.0]
    for it in range(1, STEPS + 1):
        ro, rd, gt = seq.sample_batch(N, frame=it % 8, generator=gen, device=dev)
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.render(ro[None], rd[None], cal_lidar_color=True, staged=False, perturb=True, dt_gamma=0.0, cuda_ray=mode,
                             **rkw)
            m = gt[None, :, 0]
            loss = (1e3 * (out["depth_lidar"] * m - gt[None, :, 2] * m).abs() + (out["image_lidar"][..., 0] - m) ** 2
                    + 10.0 * (out["image_lidar"][..., 1] * m - gt[None, :, 1] * m) ** 2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        c = comps(out["depth_lidar"][0].detach(), out["image_lidar"][0].detach(), gt)
        acc = [a + b for a, b in zip(acc, c)]
        if it % 100 == 0:
            img = out["image_lidar"][0].detach().float()
            print(f"     [{mode}] raydrop pred min/mean/max {float(img[:, 0].min()):.4f}/{float(img[:, 0].mean()):.4f}/{float(img[:, 0].max()):.4f} "
                  f"intensity pred min/mean/max {float(img[:, 1].min()):.4f}/{float(img[:, 1].mean()):.4f}/{float(img[:, 1].max()):.4f} "
                  f"ws mean {float(out['weights_sum_lidar'].mean()):.3f}  |g_head| {float(net.lidar_color_net.weights.grad.norm()):.3e}")
            print(f"  b2 {mode} (torch Adam, GradScaler {scaler.get_scale():.0f}) step {it:5d}: depth L1 {acc[0] / 100:.3f} m  raydrop MSE {acc[1] / 100:.4f}  "
                  f"intensity MSE {acc[2] / 100:.4f}  |w_head| {float(net.lidar_color_net.weights.norm()):.3f}", flush=True)
            acc = [0.0, 0.0, 0.0]


which = sys.argv[2] if len(sys.argv) > 2 else "all"
if which in ("all", "engine"):
    run_engine(grid_update_interval=0)
if which in ("all", "scale"):
    run_engine(grid_update_interval=0, loss_scale=4096.0)
if which in ("all", "b2"):
    run_b2()
if which in ("modes",):
    run_b2("fused")
    run_b2("ops", force_all_rays=True)
    run_b2("dense", num_steps=768, upsample_steps=64)
