"""Loss parity at matched quality (BASELINE.json: ">= 10x ... at matched depth/intensity loss").

Trains BASELINE config 2 (synthetic 64x1024 sequence, hash grid L16 F2 T2^19 + two 64x2 MLPs, 4096 rays/step, Adam
lr 1e-2) on the SAME frames with the SAME loss in three ways and evaluates each on HELD-OUT frames of the sequence with the
renderer it was trained with - the dense variant with the reference-semantic `NeRFRenderer.run` (768 + 64 samples, no
jitter; renderer.py:99-298), the engine with its occupancy march (no jitter); the engine's parameters are ALSO put through
the dense evaluator (`dense_evaluator`), which ignores the occupancy grid and therefore sees the never-trained space
behind / between the occupied cells:

  (i)   dense      the reference's sampling: render() -> dense `run` 768 + 64 (per-op sm_100a kernels + torch autograd,
                   GradScaler, torch Adam) - what the unmodified reference trains with (configs/kitti360_1908.txt:9-10);
  (ii)  ref-cuda   the occupancy-march step wired from the UNMODIFIED reference CUDA kernels (oracle/ref_cuda_step.py;
                   static LiDAR occupancy prior);
  (iii) engine     the fused engine, at max_steps 1024 / 2048 / 4096 (march step 31 / 15 / 8 cm).

Reports depth L1 (m), intensity MSE and ray-drop MSE at steps 500 / 1000 / 2000 and each variant's training rays/s ->
profiles/loss_parity.json (+ a one-screen table on stdout).
    python scripts/loss_parity.py [--steps 2000] [--skip-dense] [--skip-refcuda]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lidar_nerf_b200.data.synthetic import SyntheticLidarSequence, lidar_directions  # noqa: E402
from lidar_nerf_b200.nerf.engine import LidarFieldEngine, FieldConfig  # noqa: E402
from lidar_nerf_b200.nerf.network_tcnn import NeRFNetwork  # noqa: E402

dev = torch.device("cuda:0")
N = 4096
TRAIN = [0, 1, 2, 4, 5, 6, 8, 9]
HELD = [3, 7]
CHECK = (500, 1000, 2000)


def make_net(seq):
    net = NeRFNetwork(encoding="hashgrid", desired_resolution=32768, log2_hashmap_size=19, n_features_per_level=2,
                      num_layers=2, hidden_dim=64, geo_feat_dim=15, bound=1, density_scale=1, min_near=seq.scale,
                      min_near_lidar=seq.scale, density_thresh=10, bg_radius=-1).to(dev)
    net.grid_update_interval = 0
    return net


@torch.no_grad()
def evaluate(net, seq, emb, w_sigma, w_head):
    """Held-out frames through the dense reference-semantic renderer with the given parameters."""
    net.encoder.embeddings.data.copy_(emb.reshape(net.encoder.embeddings.shape))
    net.sigma_net.weights.data.copy_(w_sigma.reshape(-1))
    net.lidar_color_net.weights.data.copy_(w_head.reshape(-1))
    net.eval()
    d_l1, i_mse, r_mse, n_valid, n_all = 0.0, 0.0, 0.0, 0.0, 0
    dirs_s = lidar_directions(seq.H, seq.W, seq.fov_up, seq.fov, dev)
    for f in HELD:
        pose = seq.poses[f]
        rd = dirs_s @ pose[:3, :3].T
        ro = pose[:3, 3].expand_as(rd).contiguous()
        gt = seq.images[f]
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.render(ro[None], rd[None], cal_lidar_color=True, staged=True, max_ray_batch=8192, cuda_ray="dense",
                             perturb=False, num_steps=768, upsample_steps=64)
        m = gt[:, 0]
        depth = out["depth_lidar"][0].float()
        img = out["image_lidar"][0].float()
        d_l1 += float(((depth - gt[:, 2]).abs() * m).sum()) / seq.scale
        i_mse += float((((img[:, 1] - gt[:, 1]) ** 2) * m).sum())
        r_mse += float(((img[:, 0] - m) ** 2).sum())
        n_valid += float(m.sum())
        n_all += m.numel()
    return {"depth_l1_m": d_l1 / n_valid, "intensity_mse": i_mse / n_valid, "raydrop_mse": r_mse / n_all}


@torch.no_grad()
def evaluate_march(eng, seq):
    """Held-out frames through the engine's OWN renderer (occupancy march without jitter, current grid)."""
    d_l1, i_mse, r_mse, n_valid, n_all = 0.0, 0.0, 0.0, 0.0, 0
    dirs_s = lidar_directions(seq.H, seq.W, seq.fov_up, seq.fov, dev)
    for f in HELD:
        pose = seq.poses[f]
        rd = (dirs_s @ pose[:3, :3].T).contiguous()
        ro = pose[:3, 3].expand_as(rd).contiguous()
        gt = seq.images[f]
        _, depth, img = eng.render(ro, rd)
        m = gt[:, 0]
        d_l1 += float(((depth - gt[:, 2]).abs() * m).sum()) / seq.scale
        i_mse += float((((img[:, 1] - gt[:, 1]) ** 2) * m).sum())
        r_mse += float(((img[:, 0] - m) ** 2).sum())
        n_valid += float(m.sum())
        n_all += m.numel()
    return {"depth_l1_m": d_l1 / n_valid, "intensity_mse": i_mse / n_valid, "raydrop_mse": r_mse / n_all}


def batches(seq, steps, seed=0):
    gen = torch.Generator().manual_seed(seed)
    for it in range(steps):
        yield seq.sample_batch(N, frame=TRAIN[it % len(TRAIN)], generator=gen, device=dev)


def loss_fn(out, gt):
    m = gt[None, :, 0]
    return (1e3 * (out["depth_lidar"] * m - gt[None, :, 2] * m).abs() + (out["image_lidar"][..., 0] - m) ** 2
            + 10.0 * (out["image_lidar"][..., 1] * m - gt[None, :, 1] * m) ** 2).mean()


def train_dense(seq, evalnet, steps):
    torch.manual_seed(0)
    net = make_net(seq)
    net.train()
    opt = torch.optim.Adam(net.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    res, t_train = {}, 0.0
    for it, (ro, rd, gt) in enumerate(batches(seq, steps), 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.render(ro[None], rd[None], cal_lidar_color=True, staged=False, cuda_ray="dense", perturb=True,
                             num_steps=768, upsample_steps=64)
            loss = loss_fn(out, gt)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        torch.cuda.synchronize()
        t_train += time.perf_counter() - t
        if it in CHECK:
            res[it] = evaluate(evalnet, seq, net.encoder.embeddings.data, net.sigma_net.weights.data,
                               net.lidar_color_net.weights.data)
            res[it]["train_loss"] = float(loss)
            net.train()
            print(f"[dense] step {it}: {res[it]}", flush=True)
    return {"metrics": res, "rays_per_s": N * steps / t_train, "samples_per_ray": 832}


def engine_for(seq, max_steps, seed=0):
    cfg = FieldConfig(max_steps=max_steps, seed=seed)
    near, far = cfg.min_near_lidar, cfg.min_near_lidar * cfg.far_factor
    per_ray = min(max_steps, int((far - near) / (2 * 3 ** 0.5 / max_steps)) + 2)
    eng = LidarFieldEngine(cfg, N, device=dev, sample_budget=N * per_ray)
    pts = torch.cat([seq.poses[f, :3, 3] + (lidar_directions(seq.H, seq.W, seq.fov_up, seq.fov, dev)[seq.images[f][:, 0] > 0]
                                            @ seq.poses[f, :3, :3].T) * seq.images[f][seq.images[f][:, 0] > 0, 2:3] for f in TRAIN])
    eng.seed_occupancy_from_points(pts)
    return eng


def split(eng):
    a, b = eng.n_table, eng.n_table + eng.n_sigma
    P = eng.P
    return P[:a], P[a:b], P[b:eng.n_params]


def train_engine(seq, evalnet, steps, max_steps):
    eng = engine_for(seq, max_steps)
    res, t_train, spr = {}, 0.0, 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it, (ro, rd, gt) in enumerate(batches(seq, steps), 1):
        eng.set_batch(ro, rd, gt)
        e0.record()
        eng.train_step(use_graph=True)
        e1.record()
        if it % 100 == 0 or it in CHECK:
            torch.cuda.synchronize()
        if it in CHECK:
            eng.flush()
            res[it] = evaluate_march(eng, seq)                           # the renderer a user of this engine gets
            res[it]["dense_evaluator"] = evaluate(evalnet, seq, *split(eng))   # (ignores the occupancy grid)
            res[it]["train_loss"] = eng.read_loss() / max(1, it - max([0] + [c for c in CHECK if c < it]))
            spr = eng.samples_last_step()[0] / N
            res[it]["samples_per_ray"] = spr
            print(f"[engine max_steps={max_steps}] step {it}: {res[it]}", flush=True)
    # throughput: a clean timed block at the end state
    torch.cuda.synchronize()
    bl = list(batches(seq, 64, seed=5))
    for ro, rd, gt in bl[:16]:
        eng.set_batch(ro, rd, gt)
        eng.train_step(use_graph=True)
    torch.cuda.synchronize()
    e0.record()
    for ro, rd, gt in bl[16:]:
        eng.set_batch(ro, rd, gt)
        eng.train_step(use_graph=True)
    eng.flush()
    e1.record()
    torch.cuda.synchronize()
    return {"metrics": res, "rays_per_s": N * 48 / (e0.elapsed_time(e1) * 1e-3), "samples_per_ray": spr,
            "march_step_m": 2 * 3 ** 0.5 / max_steps / seq.scale}


def train_refcuda(seq, evalnet, steps):
    from oracle.ref_cuda_step import RefCudaStep
    eng = engine_for(seq, 1024)
    eng.M = N * 128
    ref = RefCudaStep(eng)
    res, t_train = {}, 0.0
    for it, (ro, rd, gt) in enumerate(batches(seq, steps), 1):
        noises = torch.rand(N, device=dev)
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = ref.step(ro, rd, gt, noises)
        torch.cuda.synchronize()
        t_train += time.perf_counter() - t
        if it in CHECK:
            res[it] = evaluate(evalnet, seq, ref.embeddings, ref.w_sigma, ref.w_head)
            res[it]["train_loss"] = float(out["loss"])
            res[it]["samples_per_ray"] = int(out["counter"][0]) / N
            print(f"[ref-cuda] step {it}: {res[it]}", flush=True)
    return {"metrics": res, "rays_per_s": N * steps / t_train}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--skip-dense", action="store_true")
    ap.add_argument("--skip-refcuda", action="store_true")
    ap.add_argument("--max-steps", type=int, nargs="*", default=[1024, 2048, 4096])
    args = ap.parse_args()
    global CHECK
    CHECK = tuple(c for c in CHECK if c <= args.steps) or (args.steps,)
    seq = SyntheticLidarSequence(n_frames=10, device=dev)
    evalnet = make_net(seq)
    out = {"workload": "BASELINE config 2, synthetic 64x1024 x 10 frames (8 train / 2 held out), 4096 rays/step, Adam lr 1e-2, "
                       "loss = 1e3 L1(depth) + MSE(ray-drop) + 10 MSE(intensity); evaluator = dense run 768+64 on held-out frames",
           "steps": args.steps, "variants": {}}
    for ms in args.max_steps:
        out["variants"][f"engine_max_steps_{ms}"] = train_engine(seq, evalnet, args.steps, ms)
    if not args.skip_refcuda:
        try:
            out["variants"]["reference_cuda_march_1024"] = train_refcuda(seq, evalnet, args.steps)
        except Exception as e:   # noqa: BLE001
            out["variants"]["reference_cuda_march_1024"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if not args.skip_dense:
        out["variants"]["dense_768_64"] = train_dense(seq, evalnet, args.steps)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "loss_parity.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
