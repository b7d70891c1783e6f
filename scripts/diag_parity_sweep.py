"""Sweep of seeds for the fused-step-vs-oracle comparison: prints the gradient errors per seed and, for outliers, where
in the hash table (which level) the disagreement sits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import check_engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
patch = "patch" in sys.argv[2:]
for seed in range(n):
    if patch:
        eng, gpu, cpu = check_engine.run_pair(n_rays=256, device="cuda:0", seed=seed, patch_smooth_gt=True,
                                              cfg=check_engine.small_config(patch_size=(2, 8), alpha_grad=100.0))
    else:
        eng, gpu, cpu = check_engine.run_pair(n_rays=256, device="cuda:0", seed=seed)
    a, b = gpu["grad"].astype(np.float64), cpu["grad"].astype(np.float64)
    nt = eng.n_table
    rel_t = np.linalg.norm(a[:nt] - b[:nt]) / np.linalg.norm(b[:nt])
    rel_w = np.linalg.norm(a[nt:] - b[nt:]) / np.linalg.norm(b[nt:])
    msg = f"seed {seed:2d}: samples {gpu['n_samples']} table rel {rel_t:.2e} mlp rel {rel_w:.2e} loss {gpu['loss']:.5f}/{cpu['loss']:.5f}"
    if rel_t > 5e-4:
        offs = eng.offsets.cpu().numpy().astype(np.int64) * eng.cfg.level_dim
        per = []
        for l in range(len(offs) - 1):
            d = np.linalg.norm(a[offs[l]:offs[l + 1]] - b[offs[l]:offs[l + 1]])
            per.append(f"{d / np.linalg.norm(b[:nt]):.1e}")
        diff = np.abs(a[:nt] - b[:nt])
        top = np.argsort(diff)[-4:]
        msg += "\n     per-level error share: " + " ".join(per) + f"\n     top elements {top.tolist()} diffs {diff[top].tolist()} ref {b[top].tolist()}"
        dws = np.abs(gpu["ws"] - cpu["ws"]).max()
        dd = np.abs(gpu["depth"] - cpu["depth"]).max()
        msg += f"\n     max |ws diff| {dws:.2e} max |depth diff| {dd:.2e}, counts equal {np.array_equal(gpu['counts'], cpu['counts'])}"
    print(msg, flush=True)
