"""The C-ABI boundary: both shared objects load and export every symbol include/b2f.h declares.
No compute calls here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b2f.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2f_[a-z0-9_]+)\s*\(", src)))


def test_header_symbol_list_matches_python_binding(ifb):
    assert _declared_symbols() == sorted(ifb._abi.SYMBOLS)


def test_product_library_exports_every_symbol(ifb):
    path = os.path.join(ROOT, "imagefiltering.jl_b200", "libb2f.so")
    assert os.path.exists(path), "libb2f.so missing: run __graft_entry__.build()"
    dll = ctypes.CDLL(path)
    for s in _declared_symbols():
        assert hasattr(dll, s), s
    lib = ifb._abi.Library(path)
    assert lib.is_device_library()
    assert "sm_100a" in lib.version()


def test_oracle_library_exports_every_symbol(ifb, oracle):
    for s in _declared_symbols():
        assert hasattr(oracle.dll, s), s
    assert not oracle.is_device_library()
    assert oracle.launch_count() == 0


def test_product_has_no_cpu_fallback(ifb):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    img = np.arange(8, dtype=np.float64)
    with pytest.raises((ifb.CudaError, ImportError)):
        ifb.imfilter(img, ifb.centered(np.ones(3) / 3))
    with pytest.raises((ifb.CudaError, ImportError)):
        ifb.mapwindow(ifb.extrema, img, 3)


def test_struct_layout_matches_header(ifb):
    # sizes implied by include/b2f.h on LP64
    assert ctypes.sizeof(ifb._abi.b2f_array) == 8 + 4 + 4 + 32 + 32 + 4 + 4
    assert ctypes.sizeof(ifb._abi.b2f_stage) == 16 + 32 + 32 + 8
    assert ctypes.sizeof(ifb._abi.b2f_border) == 8 + 8 + 32 + 32
