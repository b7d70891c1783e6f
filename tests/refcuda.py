"""Loader for the UNMODIFIED reference CUDA extensions compiled by oracle/build_ref.py into oracle/_ref/
(test infrastructure: GPU-side oracle).

On a box WITH a GPU a missing module is a hard error (every "vs reference CUDA" comparison would otherwise pass
vacuously); `LNB_ALLOW_MISSING_REF=1` downgrades it to `None` for ad-hoc runs.  Without a GPU (build container) `None`
is returned for a module that was not built."""
import importlib.machinery
import importlib.util
import os

_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
_cache = {}


def load(name):
    if name in _cache:
        return _cache[name]
    path = os.path.join(_ROOT, name, name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the extension links against libtorch)
        loader = importlib.machinery.ExtensionFileLoader(name, path)
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    if mod is None and os.environ.get("LNB_ALLOW_MISSING_REF") != "1":
        import torch
        if torch.cuda.is_available():
            raise RuntimeError(f"oracle/_ref/{name}/{name}.so is missing on a GPU box: build it in the build container "
                               "(`python oracle/build_ref.py`) so that it travels with the snapshot, or set "
                               "LNB_ALLOW_MISSING_REF=1 to skip the comparisons against the reference CUDA kernels")
    _cache[name] = mod
    return mod


def available():
    return [n for n in ("_raymarching", "_gridencoder", "_freqencoder", "_shencoder", "_ffmlp")
            if os.path.exists(os.path.join(_ROOT, n, n + ".so"))]
