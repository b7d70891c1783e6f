#!/usr/bin/env python3
"""Fill the @PLACEHOLDER@ numbers of DESIGN.md section 5 from a bench.py JSON line (gpurun_out/bench.log)."""
import json, sys, re
b = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
ref_cuda = float(sys.argv[2]) if len(sys.argv) > 2 else 219354.0
k = b["kernels_us"]
g = lambda n: f"{k[n]['us_per_step']:.0f}"
gf = k["lnb_grid_encode_forward_ex"]
rep = {
    "VALUE": f"{b['value'] / 1e6:.2f}", "MS": f"{b['ms_per_step']:.3f}", "SAMPLES": f"{b['config']['samples_per_step'] / 1e3:.0f}",
    "SPR": f"{b['config']['samples_per_ray']:.0f}", "E2E": f"{b['e2e']['value'] / 1e6:.2f}", "XREF": f"{b['value'] / ref_cuda:.0f}",
    "K_GF": g("lnb_grid_encode_forward_ex"), "K_FF": g("lnb_field_forward"), "K_CP": g("lnb_lidar_composite_step"),
    "K_HB": g("lnb_field_head_backward"), "K_SB": g("lnb_ffmlp_backward_accumulate"), "K_GB": g("lnb_grid_encode_backward_ex"),
    "K_AD": g("lnb_adam_step"), "K_MR": g("lnb_march_rays_train_ex"),
    "GF_GBS": f"{gf['algorithmic_GBps'] / 1e3:.2f}", "GF_FRAC": f"{gf['frac_of_hbm_peak']:.2f}",
}
s = open("DESIGN.md").read()
for a, v in rep.items():
    s = s.replace(f"@{a}@", v)
open("DESIGN.md", "w").write(s)
print(rep)
