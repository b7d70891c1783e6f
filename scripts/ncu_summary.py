#!/usr/bin/env python3
"""One line per captured launch of an `ncu --set full` report -> profiles/<name>.csv, and the per-launch DRAM bytes
bench.py quotes as `roofline.traffic` -> profiles/ncu_dram_bytes_per_launch.json.
    python scripts/ncu_summary.py gpurun_out/full.ncu-rep profiles/r02_ncu_full_summary.csv"""
import csv, io, json, os, re, subprocess, sys

rep, out_csv = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    v = float(r[i].replace(",", ""))
    u = units[i]
    if name.startswith("dram__bytes"):          # normalise to bytes
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
    return v


tensor = next((h for h in hdr if h.startswith("sm__pipe_tensor") and h.endswith("cycles_active.avg.pct_of_peak_sustained_active")), None)
lines = []
for r in data:
    name = re.sub(r"\(lnb::.*$|\(.*$", "", r[col["Kernel Name"]])          # drop the argument list
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"^.*?(k_[A-Za-z0-9_]+)", r"\1", name)                    # drop namespaces
    lines.append(dict(kernel=name, duration_us=round(get(r, "gpu__time_duration.sum"), 1),
                      dram_read_MB=round(get(r, "dram__bytes_read.sum") / 1e6, 1), dram_write_MB=round(get(r, "dram__bytes_write.sum") / 1e6, 1),
                      dram_pct=round(get(r, "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"), 1),
                      lts_pct=round(get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                      l1tex_pct=round(get(r, "l1tex__throughput.avg.pct_of_peak_sustained_active"), 1),
                      issue_pct=round(get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                      tensor_pipe_pct=round(get(r, tensor), 1) if tensor else 0.0,
                      warps_active_pct=round(get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                      regs=int(get(r, "launch__registers_per_thread")), grid=int(get(r, "launch__grid_size")),
                      block=int(get(r, "launch__block_size"))))
with open(out_csv, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(lines[0].keys()))
    w.writeheader()
    w.writerows(lines)
print(open(out_csv).read())

ENTRY = {"k_march_train": "lnb_march_rays_train_ex", "k_field_fused_fwd": "lnb_field_fused_forward", "k_lidar_composite_step": "lnb_lidar_composite_step",
         "k_mlp_bwd<1": "lnb_field_head_backward_rows", "k_mlp_bwd<0": "lnb_ffmlp_backward_accumulate_rows", "k_grid_bwd": "lnb_grid_encode_backward_rows",
         "k_adam": "lnb_adam_step", "k_ray_dir_terms": "lnb_field_ray_terms", "k_pack_field_weights": "lnb_field_pack_weights",
         "k_grid_fwd": "lnb_grid_encode_forward_ex", "k_field_fwd": "lnb_field_forward"}
jpath = os.path.join(os.path.dirname(out_csv), "ncu_dram_bytes_per_launch.json")
table = json.load(open(jpath)) if os.path.exists(jpath) else {}
seen = set()
for ln in lines:
    for prefix, entry in ENTRY.items():
        if ln["kernel"].startswith(prefix) and entry not in seen:
            seen.add(entry)
            table[entry] = {"dram_bytes": int(round((ln["dram_read_MB"] + ln["dram_write_MB"]) * 1e6)), "ncu_duration_us": ln["duration_us"],
                            "kernel": ln["kernel"], "source": f"{out_csv} (ncu --set full --clock-control none, steady state of bench.py)"}
json.dump(table, open(jpath, "w"), indent=1)
