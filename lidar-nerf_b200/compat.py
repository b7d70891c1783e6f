"""Make the UNMODIFIED reference code run on this library.

`install()` registers
  * the five pybind-shaped backends under the names the reference's wrappers import (`_raymarching`, `_gridencoder`,
    `_grid_encoder`, `_freqencoder`, `_shencoder`, `_sh_encoder`, `_ffmlp`; e.g. raymarching.py:5-8), and
  * this package's B2 modules under the top-level names the reference's factory imports
    (`from gridencoder import GridEncoder`, `from freqencoder import FreqEncoder`, `from shencoder import SHEncoder`,
    lidarnerf/encoding.py:68,73,78; plus `ffmlp` and `raymarching`),
  * this package's tinycudann-free `NeRFNetwork` as `lidarnerf.nerf.network_tcnn` (imported by `main_lidarnerf.py --tcnn`),
  * with `networks=True` (default) this package's `NeRFNetwork` as `lidarnerf.nerf.network` too: the reference's own
    network class renders with the dense `run` only (its renderer has no occupancy path), so
    this is what puts the sm_100a march / fused field kernels behind the unmodified `main_lidarnerf.py` and
    `Trainer` (`render()` picks the fused path on CUDA, nerf/renderer.py); `networks=False` keeps the reference's class
    (its encoders / raymarching calls still resolve to this library),
and, if the reference package `lidarnerf` is importable, fills the (empty) `lidarnerf.raymarching` namespace
(lidarnerf/raymarching/__init__.py is 0 bytes while nerf/renderer.py:140 calls raymarching.near_far_from_aabb).
"""
import importlib
import sys


def install(patch_lidarnerf=True, networks=True):
    from . import backend
    backend.install_reference_backends()
    pkg = __name__.rsplit(".", 1)[0]
    for name in ("gridencoder", "freqencoder", "shencoder", "ffmlp", "raymarching"):
        sys.modules[name] = importlib.import_module(f"{pkg}.{name}")
    # `--tcnn` / `-L` of the entry script imports lidarnerf.nerf.network_tcnn, whose reference version needs the
    # uninstallable tinycudann: pre-register this library's class under that module path (main_lidarnerf.py:289-308)
    sys.modules.setdefault("lidarnerf.nerf.network_tcnn", importlib.import_module(f"{pkg}.nerf.network_tcnn"))
    if networks:
        sys.modules.setdefault("lidarnerf.nerf.network", importlib.import_module(f"{pkg}.nerf.network"))
    # evaluation helpers the Trainer imports at module level (nerf/utils.py:23-24): the reference's chamfer module JIT-
    # compiles a CUDA extension at import time (and uses importlib.find_loader, removed in Python 3.12) - serve
    # `extern.chamfer3D.dist_chamfer_3D` / `extern.fscore` from this library's mirrors instead
    import types
    ext = importlib.import_module(f"{pkg}.extern")
    if "extern" not in sys.modules:
        root = types.ModuleType("extern")
        root.__path__ = []
        sys.modules["extern"] = root
        ch = types.ModuleType("extern.chamfer3D")
        ch.__path__ = []
        sys.modules["extern.chamfer3D"] = ch
        root.chamfer3D = ch
    sys.modules.setdefault("extern.chamfer3D.dist_chamfer_3D", importlib.import_module(f"{pkg}.extern.chamfer3D"))
    sys.modules.setdefault("extern.fscore", importlib.import_module(f"{pkg}.extern.fscore"))
    del ext
    if patch_lidarnerf:
        try:
            ref_rm = importlib.import_module("lidarnerf.raymarching")
        except Exception:
            return
        ours = sys.modules["raymarching"]
        for attr in ours.__all__:
            if not hasattr(ref_rm, attr):
                setattr(ref_rm, attr, getattr(ours, attr))


def run_script(path, argv=()):
    """Run an UNMODIFIED reference entry script (e.g. main_lidarnerf.py) on this library:
    `python -m lidar_nerf_b200.compat /path/to/main_lidarnerf.py --config configs/kitti360_1908.txt -L ...`."""
    import os
    import runpy
    install()
    root = os.path.dirname(os.path.abspath(path))
    if root not in sys.path:
        sys.path.insert(0, root)               # the script imports `lidarnerf.*` / `extern.*` relative to its own directory
    old = sys.argv
    sys.argv = [path, *argv]
    try:
        return runpy.run_path(path, run_name="__main__")
    finally:
        sys.argv = old


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m lidar_nerf_b200.compat <reference script.py> [script arguments]")
    run_script(sys.argv[1], sys.argv[2:])
