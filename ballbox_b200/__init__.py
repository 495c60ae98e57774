"""ballbox_b200 -- B200-native Ballbox PhysicsStep.

The product is ``libballbox_b200.so`` (C host library + sm_100a CUDA kernels,
built in-tree by ``ballbox_b200._build``).  This package is the Python mirror of
the reference's C API used by the tests and the benchmark; it loads that shared
object through ctypes and fails loudly when it is missing -- there is no CPU
fallback anywhere in the package.
"""
from __future__ import annotations

import os

from . import binding
from .binding import (BallboxError, BallboxLib, World, SCENE_README, SCENE_BALLS, SCENE_MIXED, SCENE_SDF,
                      SCENE_RLWORLD, SDF_NONE, SDF_PLANE_Y, SDF_WAVY_DET, SDF_WAVY_LIBM, SYNC_COMPAT,
                      SYNC_RESIDENT, DL_BODIES, DL_CONSTRAINTS)

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libballbox_b200.so")
_lib = None


def load(build_if_missing: bool = True) -> BallboxLib:
    """Load (building first if necessary) the product library."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise BallboxError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
            from . import _build
            _build.build_product()
        _lib = BallboxLib(LIB_PATH, product=True)
    return _lib


def device_count() -> int:
    return load().dll.BallboxDeviceCount()
