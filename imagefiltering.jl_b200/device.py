"""Device-resident arrays for the zero-copy form of the calls (Julia order: dims[0] fastest)."""
from __future__ import annotations

from . import _abi

_TORCH_DT = None


def _torch_dtypes():
    global _TORCH_DT
    if _TORCH_DT is None:
        import torch
        _TORCH_DT = {torch.uint8: _abi.U8, torch.int16: _abi.I16, torch.int32: _abi.I32,
                     torch.int64: _abi.I64, torch.float32: _abi.F32, torch.float64: _abi.F64}
    return _TORCH_DT


class DeviceArray:
    """A dense column-major array living in HBM.  `dims` are in Julia order (dims[0] fastest), so a
    C-contiguous torch tensor of shape (B, H, W) is the Julia array of dims (W, H, B)."""
    __slots__ = ("ptr", "dtype", "dims", "origin", "owner")

    def __init__(self, ptr, dtype, dims, origin=None, owner=None):
        self.ptr, self.dtype, self.dims = int(ptr), int(dtype), tuple(int(d) for d in dims)
        self.origin = tuple(origin) if origin is not None else (1,) * len(self.dims)
        self.owner = owner

    @classmethod
    def from_torch(cls, t, n0f8=False, origin=None):
        if not t.is_cuda:
            raise ValueError("DeviceArray.from_torch needs a CUDA tensor")
        if not t.is_contiguous():
            raise ValueError("tensor must be contiguous")
        dt = _torch_dtypes()[t.dtype]
        if n0f8:
            if dt != _abi.U8:
                raise ValueError("n0f8 needs a uint8 tensor")
            dt = _abi.N0F8
        return cls(t.data_ptr(), dt, tuple(reversed(t.shape)), origin, owner=t)

    @property
    def ndim(self):
        return len(self.dims)

    def desc(self):
        return _abi.make_array(self.ptr, self.dtype, self.dims, self.origin, _abi.DEVICE)
