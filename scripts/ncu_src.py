#!/usr/bin/env python3
"""Per-source-line stall samples of one kernel from an ncu report captured with --import-source on.
    python scripts/ncu_src.py gpurun_out/full.ncu-rep k_grid_bwd [top_n] [launch_index]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}"]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
else:
    cmd += ["--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
def _i(v):
    try:
        return int(v or 0)
    except (TypeError, ValueError):
        return 0
rows = list(csv.reader(io.StringIO(out)))
lines, cur, hdr = [], None, None
stall_cols = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur and r[0] != "":
        extra = len(r) - len(hdr)
        if extra > 0:      # unescaped quotes in the source text split the field
            r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)          # first "Source" column = CUDA text
        lines.append((cur, d))
tot = sum(_i(d.get("# Samples")) for _, d in lines)
print(f"total samples {tot}")
lines.sort(key=lambda fd: -_i(fd[1].get("# Samples")))
stalls = [k for k in (hdr or []) if k.startswith("stall_") and "Not Issued" not in k]
for f, d in lines[:top]:
    s_ = _i(d.get("# Samples"))
    st = sorted(((_i(d.get(k)), k[6:]) for k in stalls), reverse=True)[:3]
    st = " ".join(f"{k}:{v}" for v, k in st if v)
    print(f"{100*s_/max(tot,1):5.1f}%  {f.split('/')[-1]}:{d['Line No']:>4} inst={d.get('Instructions Executed','')}  [{st}]  {d['Source'].strip()[:90]}")
