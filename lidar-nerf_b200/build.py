"""Build liblnb200.so (the C-ABI CUDA library of include/lidarnerf_b200.h) in-tree with nvcc for sm_100a.

    python lidar-nerf_b200/build.py [--force] [--verbose]

One translation unit per reference extension (csrc/*.cu), compiled in parallel with
`-gencode arch=compute_100a,code=sm_100a -lineinfo`, linked into lidar-nerf_b200/lib/liblnb200.so.
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(LIB_DIR, "liblnb200.so")
UNITS = ["raymarching", "gridencoder", "freqencoder", "shencoder", "ffmlp", "field", "field_fused", "optim", "fused", "pointcloud"]
# the tensor-core MLP units are compiled a second time with bf16 operands (-DLNB_BF16, entry points `*_bf16`)
BF16_UNITS = ["ffmlp", "field", "field_fused"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; lidar-nerf_b200 has no CPU fallback and cannot be built without CUDA")


def _newest_header():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "lidarnerf_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False, trace=False):
    """trace=True builds the diagnostic variant lib/liblnb200_trace.so (-DLNB_TRACE: in-kernel timelines of the MLP
    backward, read by scripts/diag_bwd_trace.py); the shipped library never contains that code."""
    global OBJ_DIR, LIB
    nvcc = _nvcc()
    if trace:
        OBJ_DIR, LIB = os.path.join(HERE, "build_trace"), os.path.join(LIB_DIR, "liblnb200_trace.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr = _newest_header()
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u + ".cu"))]
    todo = []
    objs_all = []
    for u in units:
        for variant in (("", []), ("_bf16", ["-DLNB_BF16"])) if u in BF16_UNITS else (("", []),):
            src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(OBJ_DIR, u + variant[0] + ".o")
            objs_all.append(obj)
            if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
                todo.append((u + variant[0], src, obj, variant[1]))

    def compile_one(item):
        u, src, obj, extra = item
        cmd = [nvcc, *ARCH, *FLAGS, *extra, *(["-DLNB_TRACE"] if trace else []), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return u, r

    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for u, r in ex.map(compile_one, todo):
                if verbose or r.returncode != 0:
                    sys.stderr.write(f"--- nvcc {u} ---\n{r.stdout}{r.stderr}\n")
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on csrc/{u}.cu")
    objs = objs_all
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([nvcc, "-shared", *ARCH, "-o", LIB, *objs], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of liblnb200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, trace="--trace" in sys.argv))
