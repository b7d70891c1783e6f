"""Make the UNMODIFIED reference code run on this library.

`install()` registers
  * the five pybind-shaped backends under the names the reference's wrappers import (`_raymarching`, `_gridencoder`,
    `_grid_encoder`, `_freqencoder`, `_shencoder`, `_sh_encoder`, `_ffmlp`; e.g. raymarching.py:5-8), and
  * this package's B2 modules under the top-level names the reference's factory imports
    (`from gridencoder import GridEncoder`, `from freqencoder import FreqEncoder`, `from shencoder import SHEncoder`,
    lidarnerf/encoding.py:68,73,78; plus `ffmlp` and `raymarching`),
  * this package's tinycudann-free `NeRFNetwork` as `lidarnerf.nerf.network_tcnn` (imported by `main_lidarnerf.py --tcnn`),
and, if the reference package `lidarnerf` is importable, fills the (empty) `lidarnerf.raymarching` namespace
(lidarnerf/raymarching/__init__.py is 0 bytes while nerf/renderer.py:140 calls raymarching.near_far_from_aabb).
"""
import importlib
import sys


def install(patch_lidarnerf=True):
    from . import backend
    backend.install_reference_backends()
    pkg = __name__.rsplit(".", 1)[0]
    for name in ("gridencoder", "freqencoder", "shencoder", "ffmlp", "raymarching"):
        sys.modules[name] = importlib.import_module(f"{pkg}.{name}")
    # `--tcnn` / `-L` of the entry script imports lidarnerf.nerf.network_tcnn, whose reference version needs the
    # uninstallable tinycudann: pre-register this library's class under that module path (main_lidarnerf.py:289-308)
    sys.modules.setdefault("lidarnerf.nerf.network_tcnn", importlib.import_module(f"{pkg}.nerf.network_tcnn"))
    if patch_lidarnerf:
        try:
            ref_rm = importlib.import_module("lidarnerf.raymarching")
        except Exception:
            return
        ours = sys.modules["raymarching"]
        for attr in ours.__all__:
            if not hasattr(ref_rm, attr):
                setattr(ref_rm, attr, getattr(ours, attr))
