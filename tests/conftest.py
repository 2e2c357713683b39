import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ifb():
    import imagefiltering_jl_b200 as m
    return m


@pytest.fixture(scope="session")
def oracle(ifb):
    """The CPU oracle, loaded through the same ctypes ABI wrapper as the product library."""
    path = os.path.join(ROOT, "oracle", "libb2f_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ifb._abi.Library(path)
    assert not lib.is_device_library()
    return lib


@pytest.fixture(scope="session")
def device(ifb):
    """The product library; only gpu-marked tests may request it."""
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from importlib import import_module
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    assert lib.is_device_library()
    return lib
