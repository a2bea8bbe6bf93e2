"""secphase_b200 -- B200 (sm_100a) implementation of Secphase's marker-mode read-group scoring.

Python host layer over the C ABI of libsecphase_b200.so (include/secphase_b200.h).  There is no
CPU implementation in this package: importing works anywhere (so that the build can be checked
on a GPU-less box), but every call that computes raises SecphaseError when the CUDA library or a
CUDA device is missing.
"""
from .api import (LIB_PATH, PinnedArray, Secphase, SecphaseError, SpParams, load_library, params_for,  # noqa: F401
                  pin_batch)

__all__ = ["Secphase", "SecphaseError", "SpParams", "params_for", "load_library", "LIB_PATH", "pin_batch",
           "PinnedArray"]
