"""Import alias: the product package lives in the directory `lidar-nerf_b200/` (hyphenated, as the repo
layout requires), which Python cannot import by name.  This stub makes `import lidar_nerf_b200` resolve
to it by pointing the package search path at that directory and executing its `__init__.py`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "lidar-nerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
