"""Debug aid: where does lnb_field_fused_forward differ from the two-kernel forward?  Prints, per saved tensor, the number
of mismatching rows and how they distribute over CTA-local tile index / tile slot."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle import check_engine
from test_gpu_engine import _run_engine_only

DEV = "cuda:0"
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
extra = dict(dir_encoding="sh", sh_degree=4) if "sh" in sys.argv[2:] else {}
out = {}
for fg in (False, True):
    cfg = check_engine.small_config(fused_gather=fg, perturb=False, log2_hashmap_size=19, desired_resolution=32768, max_steps=1024,
                                    **extra)
    eng = _run_engine_only(cfg, n_rays)
    n = int(eng.counter[0])
    rays = eng.rays.cpu().numpy()
    rays = rays[np.argsort(rays[:, 0])]
    order = torch.from_numpy(np.concatenate([np.arange(o, o + k) for _, o, k in rays])).to(DEV)
    out[fg] = dict(order=order, enc=eng.enc[order], sigma=eng.sigma[order], rgb=eng.rgb[order], sig_out=eng.sig_out[order],
                   fb_s0=eng.fb_sigma[0, order], fb_s1=eng.fb_sigma[1, order], fb_h0=eng.fb_head[0, order], fb_h1=eng.fb_head[1, order])
    print(f"fused_gather={fg}: samples={n} tiles={(n + 127) // 128}")
a, b = out[True], out[False]
rows = a["order"].cpu().numpy()          # physical row of each canonical sample in the fused run
tile = rows // 128
grid = 148
k_local = tile // grid
for key in ("enc", "fb_s0", "fb_s1", "sig_out", "sigma", "fb_h0", "fb_h1", "rgb"):
    x, y = a[key].float(), b[key].float()
    bad = (x != y)
    bad = bad.reshape(bad.shape[0], -1).any(-1).cpu().numpy()
    nanx = torch.isnan(x).reshape(x.shape[0], -1).any(-1).sum().item()
    print(f"{key:8s}: {bad.sum():7d} / {len(bad)} rows differ; NaN rows in fused: {nanx}")
    if bad.any():
        for kl in range(int(k_local.max()) + 1):
            sel = k_local == kl
            if sel.any():
                print(f"      cta-local tile {kl} (slot {kl % 4}): {bad[sel].sum():6d} / {sel.sum():6d} bad")
        idx = np.nonzero(bad)[0][:3]
        for i in idx:
            print("      e.g. row", rows[i], "fused", x[i].reshape(-1)[:6].tolist(), "ref", y[i].reshape(-1)[:6].tolist())
