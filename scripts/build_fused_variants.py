"""Diagnostic builds of the persistent forward kernel: lib/liblnb200_<tag>.so = the shipped objects with csrc/field_fused.cu
recompiled under extra -D flags (selected at run time with LNB200_LIB=...; scripts/diag_fused_fwd.py times them).
    python scripts/build_fused_variants.py tag=-DFLAG[,-DFLAG2] ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lidar-nerf_b200"))
import build as B

B.build()
nvcc = B._nvcc()
for spec in sys.argv[1:]:
    tag, flags = spec.split("=", 1)
    flags = [f for f in flags.split(",") if f]
    objs = []
    for suffix, extra in (("", []), ("_bf16", ["-DLNB_BF16"])):
        obj = os.path.join(B.OBJ_DIR, f"field_fused{suffix}.{tag}.o")
        subprocess.run([nvcc, *B.ARCH, *B.FLAGS, *extra, *flags, "-c", os.path.join(B.CSRC, "field_fused.cu"), "-o", obj], check=True)
        objs.append(obj)
    others = [os.path.join(B.OBJ_DIR, f) for f in sorted(os.listdir(B.OBJ_DIR))
              if f.endswith(".o") and f.count(".") == 1 and not f.startswith("field_fused")]
    out = os.path.join(B.LIB_DIR, f"liblnb200_{tag}.so")
    subprocess.run([nvcc, "-shared", *B.ARCH, "-o", out, *others, *objs], check=True)
    print(out)
