"""B200-native colorlut / hsvfilter / hsvdetector — Python view for tests and bench.

The product is libb200vf.so (CUDA sm_100a kernels behind the C ABI of
include/b200vf.h) plus the C++ element layer in elements/; this package only
binds them.  Importable as `gst_plugins_rs_b200` (the directory name carries a
hyphen, so a one-file alias package of that name forwards here).
"""
from . import _lib, api, elements, frames, sharding  # noqa: F401
from .api import (B200VFError, Context, Group, HsvDetectorParams, HsvFilterParams,  # noqa: F401
                  frame_of, parse_cube, parse_cube_file)

__all__ = ["api", "elements", "frames", "Context", "Group", "B200VFError", "HsvFilterParams", "HsvDetectorParams",
           "frame_of", "parse_cube", "parse_cube_file"]
