"""lidar-nerf_b200: the per-ray volume-rendering hot path of tangtaogo/lidar-nerf, rebuilt for NVIDIA B200.

Layers:
  include/lidarnerf_b200.h + csrc/*.cu -> lib/liblnb200.so   C ABI, hand-written sm_100a kernels
  backend.py                                                 boundary B1: the reference's pybind modules
  raymarching/ gridencoder/ freqencoder/ shencoder/ ffmlp/   boundary B2: the reference's Python wrappers
  nerf/                                                      run_cuda glue + fused training engine
There is no CPU fallback: importing the compute modules without lib/liblnb200.so raises ImportError.
"""
__version__ = "0.1.0"

import importlib as _importlib

_LAZY = ("backend", "raymarching", "gridencoder", "freqencoder", "shencoder", "ffmlp", "activation", "encoding",
         "nerf", "data", "compat", "_lib")


def __getattr__(name):
    if name in _LAZY:
        return _importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
