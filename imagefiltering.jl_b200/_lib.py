"""Loader of the product library `libb2f.so` (CUDA sm_100a).  There is NO CPU fallback: if the
extension is missing, or no CUDA device is usable, calls fail loudly."""
from __future__ import annotations

import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2F_LIB_PATH") or os.path.join(_HERE, "libb2f.so")     # env: A/B testing of kernel builds
_lib = None


def lib() -> _abi.Library:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU execution path.")
        L = _abi.Library(LIB_PATH)
        if not L.is_device_library():
            raise ImportError(f"{LIB_PATH} is not the CUDA library")
        _lib = L
    return _lib
