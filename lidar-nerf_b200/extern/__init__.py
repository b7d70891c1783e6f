"""Mirrors of the reference's `extern/` evaluation helpers over liblnb200.so (chamfer3D, fscore)."""
from .chamfer3D import chamfer_3DDist, chamfer_3DFunction   # noqa: F401
from .fscore import fscore                                   # noqa: F401
