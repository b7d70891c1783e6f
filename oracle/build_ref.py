#!/usr/bin/env python3
"""Compile the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

This is the recipe the task allows for a compilable reference: the five in-tree extensions of
tangtaogo/lidar-nerf (`lidarnerf/{raymarching,gridencoder,freqencoder,shencoder,ffmlp}/src`) are compiled
from the sources WHERE THEY LIE under /root/reference (nothing is copied into this repo) into
`oracle/_ref/<name>/<name>.so` for sm_100a.  `oracle/_ref/` is git-ignored but NOT gpurun-ignored, so the
built modules travel to the GPU box, where `tests/test_vs_reference_cuda.py` compares our kernels against
them bit-for-bit (march sample counts) / within 1e-4 (fp32 outputs) and `tests/golden/make_golden_gpu.py`
freezes their outputs into fixtures.  Nothing in the product path (`lidar-nerf_b200/`) imports this.

Deviations from the reference's own `backend.py` build flags:
  * `-std=c++17` instead of `-std=c++14` (torch 2.11 headers refuse C++14);
  * explicit `-gencode arch=compute_100a,code=sm_100a`;
  * CUTLASS headers for ffmlp come from the image (flashinfer's vendored tree) because
    `ffmlp/dependencies/cutlass` is an un-vendored submodule in the reference.

Usage:  python oracle/build_ref.py [name ...]      (default: all five; skips ones already built)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LNB_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

EXTS = {
    # module name -> (subdir, sources)
    "_raymarching": ("raymarching", ["raymarching.cu", "bindings.cpp"]),
    "_gridencoder": ("gridencoder", ["gridencoder.cu", "bindings.cpp"]),
    "_freqencoder": ("freqencoder", ["freqencoder.cu", "bindings.cpp"]),
    "_shencoder": ("shencoder", ["shencoder.cu", "bindings.cpp"]),
    "_ffmlp": ("ffmlp", ["ffmlp.cu", "bindings.cpp"]),
    # evaluation-side extension (extern/chamfer3D): absolute source dir instead of lidarnerf/<sub>/src
    "chamfer_3D": ("@extern/chamfer3D", ["chamfer3D.cu", "chamfer_cuda.cpp"]),
}


def cutlass_includes():
    import importlib.util

    incs = []
    for pkg, rel in (("flashinfer", "data/cutlass"), ("tilelang", "3rdparty/cutlass")):
        spec = importlib.util.find_spec(pkg)
        if spec is None or not spec.submodule_search_locations:
            continue
        root = os.path.join(list(spec.submodule_search_locations)[0], rel)
        if os.path.isdir(os.path.join(root, "include", "cutlass")):
            incs = [os.path.join(root, "include"), os.path.join(root, "tools", "util", "include")]
            break
    return incs


def build(name):
    from torch.utils.cpp_extension import load

    sub, srcs = EXTS[name]
    src_dir = os.path.join(REF, sub[1:]) if sub.startswith("@") else os.path.join(REF, "lidarnerf", sub, "src")
    if not os.path.isdir(src_dir):
        print(f"[build_ref] {src_dir} not present (GPU box?) - skipping {name}")
        return False
    bdir = os.path.join(OUT, name)
    os.makedirs(bdir, exist_ok=True)
    if os.path.exists(os.path.join(bdir, name + ".so")):
        print(f"[build_ref] {name}: already built")
        return True
    cuda_flags = [
        "-O3", "-std=c++17",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
    ]
    inc = []
    if name == "_ffmlp":
        cuda_flags += ["--expt-extended-lambda", "--expt-relaxed-constexpr",
                       "-Xcompiler=-mf16c", "-Xcompiler=-Wno-float-conversion",
                       "-Xcompiler=-fno-strict-aliasing"]
        inc = cutlass_includes()
        if not inc:
            print("[build_ref] no CUTLASS headers found in the image; _ffmlp unbuildable")
            return False
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=name, sources=[os.path.join(src_dir, s) for s in srcs],
         extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=cuda_flags,
         extra_include_paths=inc, build_directory=bdir, verbose=False, is_python_module=False)
    ok = os.path.exists(os.path.join(bdir, name + ".so"))
    print(f"[build_ref] {name}: {'ok' if ok else 'FAILED'}")
    return ok


PYREF = os.path.join(OUT, "pyref")
PY_TREES = ["lidarnerf", "extern"]          # packages the entry script imports (pure Python; .cu/.cpp are not touched here)
PY_FILES = ["main_lidarnerf.py"]


def build_pyref():
    """Byte-compile the reference's UNMODIFIED Python (Trainer, datasets, entry script) from where it lies into
    SOURCELESS .pyc files under oracle/_ref/pyref/ - a binary artefact with the same status as the extension .so files
    above: git-ignored, travels to the GPU box, test infrastructure only.  `tests/test_reference_trainer.py` puts that
    directory on sys.path so the reference's own `Trainer.train_step` / `main_lidarnerf.main()` drive this library's
    kernels through `lidar_nerf_b200.compat.install()`.  No reference source text enters the repository."""
    import py_compile
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present (GPU box?) - skipping pyref")
        return False
    n = 0
    todo = [(os.path.join(REF, f), os.path.join(PYREF, f + "c")) for f in PY_FILES]
    for tree in PY_TREES:
        for root, _dirs, files in os.walk(os.path.join(REF, tree)):
            for f in files:
                if f.endswith(".py"):
                    src = os.path.join(root, f)
                    todo.append((src, os.path.join(PYREF, os.path.relpath(src, REF) + "c")))
    for src, dst in todo:
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: tracebacks name the reference path, not a path inside this repo
        py_compile.compile(src, cfile=dst, dfile=os.path.relpath(src, REF), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        n += 1
    print(f"[build_ref] pyref: {len(todo)} modules ({n} compiled now) -> {PYREF}")
    return True


if __name__ == "__main__":
    names = sys.argv[1:] or (list(EXTS) + ["pyref"])
    res = {n: (build_pyref() if n == "pyref" else build(n)) for n in names}
    sys.exit(0 if all(res.values()) else 1)
