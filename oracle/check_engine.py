"""Parity check of the fused GPU training step (lidar-nerf_b200/nerf/engine.py) against the CPU restatement
(oracle/field_step.py) on identical seeded inputs.  TEST INFRASTRUCTURE (tests/, smoke(), bench.py only)."""
import numpy as np
import torch

from . import field_step as fs


def small_config(**over):
    from lidar_nerf_b200.nerf.engine import FieldConfig
    kw = dict(num_levels=16, log2_hashmap_size=14, desired_resolution=2048, max_steps=256, loss_scale=128.0,
              grid_update_interval=0, min_near_lidar=0.02, seed=3)
    kw.update(over)
    return FieldConfig(**kw)


def run_pair(n_rays=256, device="cuda:0", cfg=None, seed=0, fill=0.2, patch_smooth_gt=False):
    """One forward/backward (no Adam) on both sides.  Returns (engine, gpu_result, cpu_result).
    patch_smooth_gt: ground-truth depth varies by < 1 cm between the rays of a patch of cfg.patch_size (so that the
    patch depth-gradient term of the loss is active on most pairs)."""
    from lidar_nerf_b200.nerf.engine import LidarFieldEngine
    cfg = cfg or small_config()
    rng = np.random.default_rng(seed)
    torch.manual_seed(4321 + seed)          # the march jitter is drawn from torch's generator: reproducible inputs
    eng = LidarFieldEngine(cfg, n_rays, device=device, sample_budget=n_rays * 96)
    # rays from a point near the origin, random directions; random ground truth
    rays_o = np.tile(rng.uniform(-0.02, 0.02, size=(1, 3)), (n_rays, 1)).astype(np.float32)
    d = rng.normal(size=(n_rays, 3))
    rays_d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    gt = np.stack([(rng.random(n_rays) < 0.9).astype(np.float32), rng.uniform(0, 1, n_rays),
                   rng.uniform(0.05, 0.8, n_rays)], -1).astype(np.float32)
    if patch_smooth_gt:
        px, py = cfg.patch_size
        n_patch = n_rays // (px * py)
        base = np.repeat(rng.uniform(0.05, 0.8, n_patch), px * py)
        wiggle = rng.uniform(-0.3, 0.3, n_patch * px * py) * 0.01 * cfg.min_near_lidar       # < 1 cm in metres
        gt[:n_patch * px * py, 2] = (base + wiggle).astype(np.float32)
    # a clumpy occupancy grid
    bits = np.zeros(cfg.cascade * cfg.grid_size ** 3, bool)
    pos = 0
    while pos < bits.size:
        run = int(rng.integers(1, 4096))
        if rng.random() < fill:
            bits[pos:pos + run] = True
        pos += run
    bitfield = np.packbits(bits.reshape(-1, 8), axis=1, bitorder="little").reshape(-1)
    # make the randomly initialised field non-trivial: larger table values
    P = eng.P.cpu().numpy().copy()
    P[:eng.n_table] = rng.uniform(-0.5, 0.5, size=eng.n_table).astype(np.float32)
    eng.P.copy_(torch.from_numpy(P))
    eng.Ph.copy_(eng.P.to(torch.float16))
    P = P[:eng.n_params]

    eng.bitfield.copy_(torch.from_numpy(bitfield))
    eng.set_batch(torch.from_numpy(rays_o).to(device), torch.from_numpy(rays_d).to(device),
                  torch.from_numpy(gt).to(device))
    eng.G.zero_()
    eng.loss_acc.zero_()
    eng._forward_backward()
    torch.cuda.synchronize()
    noises = eng.noises.cpu().numpy()

    params = fs.FieldParams(cfg, P)
    L = cfg.num_levels
    ls = (torch.exp2(torch.arange(L, device=device, dtype=torch.float32) * torch.tensor(eng.S, device=device))
          * cfg.base_resolution - 1.0).cpu().numpy()
    cpu = fs.field_step(params, rays_o, rays_d, gt, noises, bitfield, eng.M, level_scales=ls, apply_adam=False)
    rays = eng.rays.cpu().numpy()
    order = np.argsort(rays[:, 0])
    gpu = dict(loss=float(eng.loss_acc.item()), grad=eng.G[:eng.n_params].cpu().numpy(), counts=rays[order, 2], ws=eng.ws.cpu().numpy(),
               depth=eng.depth.cpu().numpy(), image=eng.image.cpu().numpy(), n_samples=int(eng.counter[0].item()))
    return eng, gpu, cpu


def _record(name, **vals):
    try:
        from conftest import record_parity      # tests/ on sys.path: only under pytest
        record_parity(name, **vals)
    except Exception:
        pass


def compare(gpu, cpu, n_table, verbose=True):
    """Raises AssertionError with a diagnostic when the fused step disagrees with the restatement."""
    np.testing.assert_array_equal(gpu["counts"], cpu["counts"], err_msg="per-ray sample counts")
    assert gpu["n_samples"] == cpu["n_samples"]
    # fp16 MLPs + fast-math exp/sin: outputs agree to ~1e-3 relative
    np.testing.assert_allclose(gpu["ws"], cpu["ws"], rtol=5e-3, atol=2e-3)
    np.testing.assert_allclose(gpu["depth"], cpu["depth"], rtol=5e-3, atol=2e-3)
    np.testing.assert_allclose(gpu["image"], cpu["image"], rtol=5e-3, atol=2e-3)
    np.testing.assert_allclose(gpu["loss"], cpu["loss"], rtol=5e-3)
    gg, cg = gpu["grad"].astype(np.float64), cpu["grad"].astype(np.float64)
    for name, sl in (("hash table", slice(0, n_table)), ("MLP weights", slice(n_table, None))):
        a, b = gg[sl], cg[sl]
        rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        cos = float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
        if verbose:
            print(f"[check_engine] grad {name}: |g|={np.linalg.norm(b):.4e} rel.err={rel:.3e} cos={cos:.6f}")
        _record(f"check_engine.compare[{name}]", rel=rel, one_minus_cos=1 - cos)
        # observed on B200: rel 4e-5 .. 3e-4 (fp16 activations / fast-math exp, fp32 accumulation on both sides)
        assert rel < 2e-3 and cos > 0.99999, f"gradient of {name} disagrees: rel={rel:.3e} cos={cos:.6f}"
