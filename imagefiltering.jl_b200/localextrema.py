"""`findlocalmaxima`, `findlocalminima`, `blob_LoG` (reference src/extrema.jl) on top of the LoG path.

Host-side mirror only: index bookkeeping and the blob list.  The LoG filtering (`imfilter!` per σ), the `-σ` scaling
into the σ-stack, the strict-peak scan, `maximum(abs, img)` and the amplitude gather all run behind the C ABI
(`b2f_imfilter`, `b2f_scale_into_slice`, `b2f_findlocalextrema`, `b2f_maxabs`, `b2f_gather`); with the CUDA library
the σ-stack never leaves the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi, kernel as Kernel
from ._abi import ArgumentError
from .border import Pad
from .device import DeviceArray
from .imfilter import _as_input, build_stages, factorkernel
from .n0f8 import N0f8Array


@dataclass(frozen=True)
class BlobLoG:
    """src/extrema.jl:1-18.  `location` is the 1-based CartesianIndex of the peak, `σ` the scale tuple that gave the
    largest -LoG amplitude there, `amplitude` that value (radius = σ√2)."""
    location: tuple
    σ: tuple
    amplitude: float

    def __repr__(self):
        return f"BlobLoG(location=CartesianIndex{self.location}, σ={self.σ}, amplitude={self.amplitude})"


def _lib(_library):
    if _library is not None:
        return _library
    from ._lib import lib
    return lib()                        # raises when the CUDA extension is missing: no CPU path


def _unravel_array(lin, shape, first=None):
    """0-based column-major linear indices -> (n, N) int64 array of indices along the array's own axes (Julia
    CartesianIndex rows: 1-based for a plain Array, offset by `first` for an OffsetArray)."""
    if len(lin) == 0:
        return np.empty((0, len(shape)), dtype=np.int64)
    base = np.asarray(first if first is not None else (1,) * len(shape), dtype=np.int64)
    return np.stack(np.unravel_index(lin, shape, order="F"), axis=1).astype(np.int64) + base


def _unravel(lin, shape, first=None):
    return list(map(tuple, _unravel_array(lin, shape, first).tolist()))


def _findlocalextrema(minima, img, window, edges, _library, as_array=False):
    desc, ndim, first, shape, keep = _as_input(img)
    if window is None:
        window = (3,) * ndim            # default_window: 3 on every spatial axis (src/extrema.jl:107)
    if isinstance(edges, bool):
        edges = (edges,) * ndim
    if len(window) != ndim or len(edges) != ndim:
        raise ArgumentError("window and edges need one entry per dimension of img")
    lin = _lib(_library).findlocalextrema(desc, minima, window, edges)
    # CartesianIndices(img) carry the array's own axes (src/extrema.jl:125-162): an OffsetArray reports offset indices
    return _unravel_array(lin, shape, first) if as_array else _unravel(lin, shape, first)


def findlocalmaxima(img, *, window=None, edges=True, as_array=False, _library=None):
    """findlocalmaxima(img; window=default_window(img), edges=true) -> Vector{CartesianIndex}
    (src/extrema.jl:107-119): the elements larger than all of their neighbours inside the window, as a list of 1-based
    index tuples in the reference's order (`as_array=True`: the same rows as an (n, N) int64 array, no Python objects)."""
    return _findlocalextrema(False, img, window, edges, _library, as_array)


def findlocalminima(img, *, window=None, edges=True, as_array=False, _library=None):
    """src/extrema.jl:121-127."""
    return _findlocalextrema(True, img, window, edges, _library, as_array)


def blob_LoG(img, σscales, *, edges=None, σshape=None, rthresh=1e-3, _library=None):
    """blob_LoG(img, σscales; edges=(true, false, ...), σshape=(1, ...), rthresh=0.001) -> Vector{BlobLoG}
    (src/extrema.jl:72-92, multiLoG :94-105)."""
    lib = _lib(_library)
    desc, N, first, shape, keep = _as_input(img)
    if N + 1 > _abi.MAXDIM:
        raise _abi.NotSupportedError("blob_LoG needs one axis more than img (at most 4 in total)")
    if edges is None:
        edges = (True,) + (False,) * N
    elif isinstance(edges, bool):
        edges = (edges,) * (N + 1)
    if len(edges) != N + 1:
        raise ArgumentError("edges needs N+1 entries: the σ axis first")
    σshape = tuple(float(s) for s in (σshape if σshape is not None else (1,) * N))
    sigmas = sorted(float(s) for s in σscales)
    S = len(sigmas)
    in_dt = desc.dtype
    F = np.float32 if in_dt in (_abi.F32, _abi.N0F8) else np.float64          # float(eltype(T))
    fdt = _abi.F32 if F == np.float32 else _abi.F64
    numel = int(np.prod(shape))
    esz = np.dtype(F).itemsize
    border = Pad("reflect").to_abi(N)
    device = lib.is_device_library()
    owned = []
    try:
        if device:                                                          # ONE device allocation for all temporaries
            in_bytes = numel * _abi.DTYPE_SIZE[in_dt] if desc.mem == _abi.HOST else 0
            in_bytes = (in_bytes + 255) // 256 * 256
            base = lib.malloc(in_bytes + (S + 1) * numel * esz)
            owned.append(base)
            if desc.mem == _abi.HOST:                                       # upload once, filter S times
                lib.check(lib.dll.b2f_memcpy_h2d(base, desc.ptr, numel * _abi.DTYPE_SIZE[in_dt]))
                desc = _abi.make_array(base, in_dt, shape, first, _abi.DEVICE)
            tmp_ptr = base + in_bytes
            stack_ptr = tmp_ptr + numel * esz
            mem = _abi.DEVICE
        else:                                                               # the oracle library works on host arrays
            tmp_np = np.empty(shape, dtype=F, order="F")
            stack_np = np.empty((S,) + tuple(shape), dtype=F, order="F")
            tmp_ptr, stack_ptr, mem = tmp_np.ctypes.data, stack_np.ctypes.data, _abi.HOST
        tmp = _abi.make_array(tmp_ptr, fdt, shape, first, mem)
        stack = _abi.make_array(stack_ptr, fdt, (S,) + tuple(shape), (1,) + tuple(first), mem)
        for i, σ in enumerate(sigmas):
            k = Kernel.LoG(tuple(σ * s for s in σshape)) if N > 1 else Kernel.LoG((σ * σshape[0],))
            stages = _abi.StageList(build_stages(factorkernel(k), N))
            lib.imfilter(desc, tmp, stages, border)                          # imfilter!(LoG_slice, img, Kernel.LoG(σ), "reflect")
            lib.scale_into_slice(tmp, stack, i, -σ)                          # LoG_slice .*= -σ
        peaks = lib.findlocalextrema(stack, False, (3,) * (N + 1), edges)    # findlocalmaxima(img_LoG; edges)
        amps = lib.gather(stack, peaks)
        imgmax = lib.maxabs(desc) if rthresh != 0 else 0.0
        if in_dt == _abi.N0F8:
            imgmax /= 255.0
    finally:
        if owned:
            lib.check(lib.dll.b2f_sync())
        for p in owned:
            lib.free(p)
    # the reference's comprehension (src/extrema.jl:86-90), thresholded on arrays before any Python object is made
    sidx = peaks % S                                                         # x[1] - 1: the σ index is the leading axis
    sig = np.asarray(sigmas)
    if rthresh != 0:
        athresh = rthresh / (sig ** N * float(np.prod(σshape)))              # src/extrema.jl:84
        keep_mask = amps > athresh[sidx] * imgmax
        peaks, amps, sidx = peaks[keep_mask], amps[keep_mask], sidx[keep_mask]
    locs = _unravel_array(peaks, (S,) + tuple(shape))[:, 1:] + (np.asarray(first, dtype=np.int64) - 1)
    blobs = []
    for loc, s, a in zip(locs.tolist(), sidx.tolist(), amps.tolist()):
        blobs.append(BlobLoG(tuple(loc), tuple(sigmas[s] * t for t in σshape), F(a).item() if F == np.float32 else a))
    return blobs
