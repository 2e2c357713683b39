"""ctypes view of include/b2f.h.

`Library(path)` wraps one shared object exporting the b2f C ABI.  The product package only ever
instantiates it on `libb2f.so` (see `_lib.py`); the test-suite instantiates a second one on
`oracle/libb2f_oracle.so`, which exports the same symbols.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAXDIM = 4
MAXSTAGES = 16

# dtypes
U8, N0F8, I16, I32, I64, F32, F64, U16, U32 = range(9)
HOST, DEVICE = 0, 1
REPLICATE, CIRCULAR, SYMMETRIC, REFLECT, FILL, INNER, NOPAD = range(7)
STAGE_1D, STAGE_DENSE = 0, 1
TAPS_F64, TAPS_F32, TAPS_INT = 0, 1, 2
WIN_MEDIAN, WIN_MEAN, WIN_SUM, WIN_MIN, WIN_MAX = range(5)
OK, EDIM, EARG, EINEXACT, ECUDA, ENOTSUP, ENOMEM = 0, -1, -2, -3, -4, -5, -6

DTYPE_SIZE = {U8: 1, N0F8: 1, I16: 2, I32: 4, I64: 8, F32: 4, F64: 8, U16: 2, U32: 4}
NP_TO_DTYPE = {
    np.dtype(np.uint8): U8, np.dtype(np.int16): I16, np.dtype(np.int32): I32,
    np.dtype(np.int64): I64, np.dtype(np.float32): F32, np.dtype(np.float64): F64,
    np.dtype(np.uint16): U16, np.dtype(np.uint32): U32,
}
DTYPE_TO_NP = {v: k for k, v in NP_TO_DTYPE.items()}
DTYPE_TO_NP[N0F8] = np.dtype(np.uint8)

# every symbol include/b2f.h declares (tests check the built library exports all of them)
ACCUM_EXACT, ACCUM_FMA = 0, 1

SYMBOLS = [
    "b2f_version", "b2f_last_error", "b2f_is_device_library", "b2f_set_device", "b2f_device_count", "b2f_sm_count",
    "b2f_malloc", "b2f_free", "b2f_host_alloc", "b2f_host_free", "b2f_host_register", "b2f_host_unregister", "b2f_memcpy_h2d", "b2f_memcpy_d2h",
    "b2f_sync", "b2f_ipc_export", "b2f_ipc_open", "b2f_ipc_close", "b2f_imfilter", "b2f_imgradients", "b2f_mapwindow_extrema", "b2f_mapwindow_median", "b2f_mapwindow_reduce", "b2f_imfilter_slab", "b2f_imfilter_slab_staged", "b2f_imfilter_slab_xy", "b2f_shard_ctx_create", "b2f_shard_ctx_export", "b2f_shard_ctx_connect", "b2f_shard_handshake", "b2f_imfilter_sharded", "b2f_shard_ctx_destroy", "b2f_memcpy_async", "b2f_memcpy2d_async",
    "b2f_memset_async", "b2f_stream_write32", "b2f_stream_wait_geq32",
    "b2f_findlocalextrema", "b2f_scale_into_slice", "b2f_maxabs", "b2f_gather", "b2f_na_prepare", "b2f_divide",
    "b2f_normalize_dims",
    "b2f_bench_fma_peak", "b2f_set_accum_mode", "b2f_iir", "b2f_imfilter_fft", "b2f_launch_count", "b2f_reset_launch_count", "b2f_last_path",
]


class b2f_array(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
        ("dims", C.c_int64 * MAXDIM), ("origin", C.c_int64 * MAXDIM),
        ("mem", C.c_int32), ("reserved", C.c_int32),
    ]


class b2f_stage(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("axis", C.c_int32), ("ndim", C.c_int32), ("tap_dtype", C.c_int32),
        ("len", C.c_int64 * MAXDIM), ("lo", C.c_int64 * MAXDIM), ("taps", C.POINTER(C.c_double)),
    ]


class b2f_border(C.Structure):
    _fields_ = [
        ("style", C.c_int32), ("npad", C.c_int32), ("fill", C.c_double),
        ("lo", C.c_int64 * MAXDIM), ("hi", C.c_int64 * MAXDIM),
    ]


class b2f_slab_xy(C.Structure):
    """xy-filtered boundary planes of a slab (include/b2f.h)."""
    _fields_ = [
        ("xy_lo", C.c_void_p), ("lo_halo", C.c_int64), ("lo_own", C.c_int64),
        ("xy_hi", C.c_void_p), ("hi_own", C.c_int64), ("hi_halo", C.c_int64),
    ]


class DimensionMismatch(ValueError):
    """Julia `DimensionMismatch` (reference src/imfilter.jl:604-615)."""


class ArgumentError(ValueError):
    """Julia `ArgumentError` (reference src/border.jl:147-155)."""


class InexactError(ArithmeticError):
    """Julia `InexactError` (reference src/imfilter.jl:233-254)."""


class CudaError(RuntimeError):
    pass


class NotSupportedError(NotImplementedError):
    """A valid reference call that this library does not accelerate (there is no CPU fallback)."""


_EXC = {EDIM: DimensionMismatch, EARG: ArgumentError, EINEXACT: InexactError, ECUDA: CudaError,
        ENOTSUP: NotSupportedError, ENOMEM: MemoryError}


def make_array(ptr: int, dtype: int, dims, origin=None, mem=HOST) -> b2f_array:
    a = b2f_array()
    a.ptr = ptr
    a.dtype = dtype
    a.ndim = len(dims)
    if len(dims) > MAXDIM:
        raise NotSupportedError(f"arrays with more than {MAXDIM} dimensions are not supported")
    for d in range(MAXDIM):
        a.dims[d] = dims[d] if d < len(dims) else 1
        a.origin[d] = (origin[d] if origin is not None else 1) if d < len(dims) else 0
    a.mem = mem
    return a


def numpy_array_desc(x: np.ndarray, origin=None, dtype=None) -> b2f_array:
    """Describe a Fortran-contiguous numpy array (axis 0 fastest == Julia dim 1)."""
    if x.ndim > 1 and not x.flags.f_contiguous:
        raise ValueError("array must be Fortran-contiguous (Julia memory order)")
    if x.ndim == 1 and not x.flags.c_contiguous:
        raise ValueError("vector must be contiguous")
    if dtype is None and x.dtype not in NP_TO_DTYPE:
        raise NotSupportedError(f"element type {x.dtype} has no counterpart in the C ABI (include/b2f.h dtypes)")
    dt = NP_TO_DTYPE[x.dtype] if dtype is None else dtype
    return make_array(x.ctypes.data, dt, x.shape, origin, HOST)


class StageList:
    """Keeps the tap buffers alive next to the ctypes stage array."""

    def __init__(self, stages):
        # stages: list of dicts {kind, axis, ndim, tap_dtype, len(list), lo(list), taps(np f64, F-order)}
        n = len(stages)
        self.n = n
        self.arr = (b2f_stage * max(n, 1))()
        self._keep = []
        for i, s in enumerate(stages):
            st = self.arr[i]
            st.kind = s["kind"]
            st.axis = s.get("axis", 0)
            st.ndim = s.get("ndim", 0)
            st.tap_dtype = s.get("tap_dtype", TAPS_F64)
            taps = np.ascontiguousarray(np.asarray(s["taps"], dtype=np.float64).ravel(order="F"))
            self._keep.append(taps)
            for d in range(MAXDIM):
                st.len[d] = s["len"][d] if d < len(s["len"]) else 1
                st.lo[d] = s["lo"][d] if d < len(s["lo"]) else 0
            st.taps = taps.ctypes.data_as(C.POINTER(C.c_double))


def make_border(style: int, fill: float = 0.0, lo=None, hi=None) -> b2f_border:
    b = b2f_border()
    b.style = style
    b.fill = float(fill)
    if lo is not None:
        b.npad = len(lo)
        for d in range(len(lo)):
            b.lo[d] = lo[d]
            b.hi[d] = hi[d]
    else:
        b.npad = 0
    return b


class Library:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise ImportError(f"b2f shared library not found: {path}")
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.b2f_version.restype = C.c_char_p
        d.b2f_last_error.restype = C.c_char_p
        d.b2f_last_path.restype = C.c_char_p
        d.b2f_launch_count.restype = C.c_int64
        d.b2f_reset_launch_count.restype = None
        d.b2f_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
        d.b2f_free.argtypes = [C.c_void_p]
        d.b2f_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
        d.b2f_host_free.argtypes = [C.c_void_p]
        d.b2f_host_register.argtypes = [C.c_void_p, C.c_uint64]
        d.b2f_host_unregister.argtypes = [C.c_void_p]
        d.b2f_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        d.b2f_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        d.b2f_set_device.argtypes = [C.c_int]
        d.b2f_set_accum_mode.argtypes = [C.c_int32]
        d.b2f_imfilter_fft.argtypes = [C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.POINTER(b2f_border),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
        d.b2f_iir.argtypes = [C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_int32, C.POINTER(C.c_double), C.POINTER(b2f_border), C.c_void_p]
        d.b2f_device_count.argtypes = [C.POINTER(C.c_int)]
        d.b2f_sm_count.argtypes = [C.POINTER(C.c_int)]
        d.b2f_ipc_export.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        d.b2f_ipc_open.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        d.b2f_ipc_close.argtypes = [C.c_void_p, C.c_uint64]
        d.b2f_imfilter.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.c_int32,
            C.POINTER(b2f_border), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
        d.b2f_imgradients.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_int32, C.POINTER(b2f_stage), C.c_int32,
            C.POINTER(b2f_border), C.c_void_p]
        d.b2f_mapwindow_extrema.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_int32,
            C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(b2f_border), C.c_void_p]
        d.b2f_mapwindow_median.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(b2f_border), C.c_void_p]
        d.b2f_mapwindow_reduce.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(b2f_border),
            C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]
        d.b2f_imfilter_slab.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.c_int32,
            C.POINTER(b2f_border), C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]

        d.b2f_imfilter_slab_staged.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.c_int32,
            C.POINTER(b2f_border), C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
            C.c_int32, C.c_int32, C.c_void_p]
        d.b2f_imfilter_slab_xy.argtypes = [
            C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.c_int32,
            C.POINTER(b2f_border), C.c_int64, C.c_int64, C.POINTER(b2f_slab_xy), C.c_void_p, C.c_void_p,
            C.c_int32, C.c_int32, C.c_void_p]
        d.b2f_shard_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_int32]
        d.b2f_shard_ctx_export.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        d.b2f_shard_ctx_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        d.b2f_shard_handshake.argtypes = [C.c_void_p, C.c_void_p]
        d.b2f_imfilter_sharded.argtypes = [C.c_void_p, C.POINTER(b2f_array), C.POINTER(b2f_array), C.POINTER(b2f_stage), C.c_int32,
                                           C.POINTER(b2f_border), C.c_int64, C.c_int64, C.c_void_p]
        d.b2f_shard_ctx_destroy.argtypes = [C.c_void_p]
        d.b2f_memcpy2d_async.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        d.b2f_memcpy_async.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        d.b2f_memset_async.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p]
        d.b2f_stream_write32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        d.b2f_stream_wait_geq32.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        d.b2f_findlocalextrema.argtypes = [
            C.POINTER(b2f_array), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int64,
            C.POINTER(C.c_int64), C.c_void_p]
        d.b2f_scale_into_slice.argtypes = [C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_int64, C.c_double, C.c_void_p]
        d.b2f_maxabs.argtypes = [C.POINTER(b2f_array), C.POINTER(C.c_double), C.c_void_p]
        d.b2f_na_prepare.argtypes = [C.POINTER(b2f_array), C.c_int32, C.POINTER(b2f_array), C.POINTER(b2f_array),
                                     C.POINTER(C.c_int32), C.c_void_p]
        d.b2f_divide.argtypes = [C.POINTER(b2f_array), C.POINTER(b2f_array), C.c_void_p]
        d.b2f_normalize_dims.argtypes = [C.POINTER(b2f_array), C.POINTER(C.POINTER(C.c_double)), C.c_void_p]
        d.b2f_gather.argtypes = [C.POINTER(b2f_array), C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_double), C.c_void_p]

    # -- helpers ---------------------------------------------------------------------------
    def check(self, rc: int):
        if rc == OK:
            return
        msg = (self.dll.b2f_last_error() or b"").decode()
        raise _EXC.get(rc, RuntimeError)(msg or f"b2f status {rc}")

    def version(self) -> str:
        return self.dll.b2f_version().decode()

    def is_device_library(self) -> bool:
        return bool(self.dll.b2f_is_device_library())

    def imfilter_fft(self, img: b2f_array, out: b2f_array, stages: "StageList", border: b2f_border, roi=None, stream: int = 0):
        """b2f_imfilter_fft: `stages` holds exactly one dense stage (the convolution of the kernel's factors)."""
        lo = hi = None
        if roi is not None:
            lo = (C.c_int64 * MAXDIM)(*list(roi[0]) + [0] * (MAXDIM - len(roi[0])))
            hi = (C.c_int64 * MAXDIM)(*list(roi[1]) + [0] * (MAXDIM - len(roi[1])))
        self.check(self.dll.b2f_imfilter_fft(C.byref(img), C.byref(out), stages.arr, C.byref(border), lo, hi, C.c_void_p(stream)))

    def iir(self, img: b2f_array, out: b2f_array, axis: int, coef, border: b2f_border, stream: int = 0):
        """b2f_iir: one Triggs-Sdika recursive pass along `axis` (0-based); coef = TriggsSdika.coefficients() (18 doubles)."""
        c = (C.c_double * 18)(*[float(x) for x in coef])
        self.check(self.dll.b2f_iir(C.byref(img), C.byref(out), int(axis), c, C.byref(border), C.c_void_p(stream)))

    def set_accum_mode(self, mode: int) -> int:
        """b2f_set_accum_mode: ACCUM_EXACT (0) / ACCUM_FMA (1) for the calling thread; returns the previous mode."""
        prev = int(self.dll.b2f_set_accum_mode(int(mode)))
        if prev < 0:
            self.check(prev)
        return prev

    def sm_count(self) -> int:
        n = C.c_int()
        self.check(self.dll.b2f_sm_count(C.byref(n)))
        return int(n.value)

    def last_path(self) -> str:
        return self.dll.b2f_last_path().decode()

    def launch_count(self) -> int:
        return int(self.dll.b2f_launch_count())

    def reset_launch_count(self):
        self.dll.b2f_reset_launch_count()

    # -- peer mapping ------------------------------------------------------------------------
    def ipc_export(self, dptr: int):
        """-> (64-byte handle, offset) naming device memory of this process for b2f_ipc_open elsewhere."""
        h = C.create_string_buffer(64)
        off = C.c_uint64()
        self.check(self.dll.b2f_ipc_export(C.c_void_p(dptr), h, C.byref(off)))
        return h.raw, int(off.value)

    def ipc_open(self, handle: bytes, offset: int) -> int:
        p = C.c_void_p()
        self.check(self.dll.b2f_ipc_open(C.create_string_buffer(handle, 64), offset, C.byref(p)))
        return int(p.value)

    def ipc_close(self, dptr: int, offset: int):
        self.check(self.dll.b2f_ipc_close(C.c_void_p(dptr), offset))

    # -- raw calls on descriptors -------------------------------------------------------------
    def imfilter(self, img: b2f_array, out: b2f_array, stages: StageList, border: b2f_border,
                 roi=None, stream: int = 0):
        if roi is not None:
            lo = (C.c_int64 * MAXDIM)(*list(roi[0]) + [0] * (MAXDIM - len(roi[0])))
            hi = (C.c_int64 * MAXDIM)(*list(roi[1]) + [0] * (MAXDIM - len(roi[1])))
        else:
            lo = hi = None
        self.check(self.dll.b2f_imfilter(C.byref(img), C.byref(out), stages.arr, stages.n,
                                         C.byref(border), lo, hi, C.c_void_p(stream)))

    def imgradients(self, img: b2f_array, outs, stages: StageList, nstages_each: int,
                    border: b2f_border, stream: int = 0):
        arr = (b2f_array * len(outs))(*outs)
        self.check(self.dll.b2f_imgradients(C.byref(img), arr, len(outs), stages.arr, nstages_each,
                                            C.byref(border), C.c_void_p(stream)))

    def mapwindow_extrema(self, img: b2f_array, out_min, out_max, interleaved: bool, win_lo, win_hi,
                          border: b2f_border, stream: int = 0):
        lo = (C.c_int64 * MAXDIM)(*list(win_lo) + [0] * (MAXDIM - len(win_lo)))
        hi = (C.c_int64 * MAXDIM)(*list(win_hi) + [0] * (MAXDIM - len(win_hi)))
        self.check(self.dll.b2f_mapwindow_extrema(
            C.byref(img), C.byref(out_min) if out_min is not None else None,
            C.byref(out_max) if out_max is not None else None, 1 if interleaved else 0, lo, hi,
            C.byref(border), C.c_void_p(stream)))

    def mapwindow_median(self, img: b2f_array, out: b2f_array, win_lo, win_hi, border: b2f_border, stream: int = 0):
        lo = (C.c_int64 * MAXDIM)(*list(win_lo) + [0] * (MAXDIM - len(win_lo)))
        hi = (C.c_int64 * MAXDIM)(*list(win_hi) + [0] * (MAXDIM - len(win_hi)))
        self.check(self.dll.b2f_mapwindow_median(C.byref(img), C.byref(out), lo, hi, C.byref(border), C.c_void_p(stream)))

    def mapwindow_reduce(self, img: b2f_array, out: b2f_array, op: int, win_lo, win_hi, border: b2f_border, idx_first=None,
                         idx_step=None, stream: int = 0):
        lo = (C.c_int64 * MAXDIM)(*list(win_lo) + [0] * (MAXDIM - len(win_lo)))
        hi = (C.c_int64 * MAXDIM)(*list(win_hi) + [0] * (MAXDIM - len(win_hi)))
        f = s = None
        if idx_first is not None:
            f = (C.c_int64 * MAXDIM)(*list(idx_first) + [0] * (MAXDIM - len(idx_first)))
            s = (C.c_int64 * MAXDIM)(*list(idx_step) + [1] * (MAXDIM - len(idx_step)))
        self.check(self.dll.b2f_mapwindow_reduce(C.byref(img), C.byref(out), op, lo, hi, C.byref(border), f, s, C.c_void_p(stream)))

    def imfilter_slab(self, img: b2f_array, out: b2f_array, stages: StageList, border: b2f_border,
                      global_last_dim: int, slab_first: int, halo_lo: int, n_halo_lo: int,
                      halo_hi: int, n_halo_hi: int, stream: int = 0):
        """halo_lo / halo_hi are raw pointers (ints; 0 = none) to n_halo_* input planes below / above the slab."""
        self.check(self.dll.b2f_imfilter_slab(C.byref(img), C.byref(out), stages.arr, stages.n,
                                              C.byref(border), global_last_dim, slab_first,
                                              C.c_void_p(halo_lo or None), n_halo_lo,
                                              C.c_void_p(halo_hi or None), n_halo_hi, C.c_void_p(stream)))

    def imfilter_slab_staged(self, img: b2f_array, out: b2f_array, stages: StageList, border: b2f_border,
                             global_last_dim: int, slab_first: int, halo_lo: int, n_halo_lo: int, halo_hi: int,
                             n_halo_hi: int, flag_lo: int, flag_hi: int, epoch: int, lo_early_rows: int = 0, stream: int = 0):
        self.check(self.dll.b2f_imfilter_slab_staged(C.byref(img), C.byref(out), stages.arr, stages.n, C.byref(border),
                                                     global_last_dim, slab_first, C.c_void_p(halo_lo or None), n_halo_lo,
                                                     C.c_void_p(halo_hi or None), n_halo_hi, C.c_void_p(flag_lo or None),
                                                     C.c_void_p(flag_hi or None), epoch, lo_early_rows, C.c_void_p(stream)))

    def imfilter_slab_xy(self, img: b2f_array, out: b2f_array, stages: StageList, border: b2f_border,
                         global_last_dim: int, slab_first: int, xy_lo: int, lo_halo: int, lo_own: int, xy_hi: int,
                         hi_own: int, hi_halo: int, flag_lo: int = 0, flag_hi: int = 0, epoch: int = 0,
                         lo_early_rows: int = 0, stream: int = 0):
        xy = b2f_slab_xy(C.c_void_p(xy_lo or None), lo_halo, lo_own, C.c_void_p(xy_hi or None), hi_own, hi_halo)
        self.check(self.dll.b2f_imfilter_slab_xy(C.byref(img), C.byref(out), stages.arr, stages.n, C.byref(border),
                                                 global_last_dim, slab_first, C.byref(xy), C.c_void_p(flag_lo or None),
                                                 C.c_void_p(flag_hi or None), epoch, lo_early_rows, C.c_void_p(stream)))

    def memcpy_async(self, dst: int, src: int, nbytes: int, stream: int = 0):
        self.check(self.dll.b2f_memcpy_async(C.c_void_p(dst), C.c_void_p(src), nbytes, C.c_void_p(stream)))

    def memcpy2d_async(self, dst: int, dpitch: int, src: int, spitch: int, width: int, height: int, stream: int = 0):
        self.check(self.dll.b2f_memcpy2d_async(C.c_void_p(dst), dpitch, C.c_void_p(src), spitch, width, height, C.c_void_p(stream)))

    def stream_write32(self, dptr: int, value: int, stream: int = 0):
        self.check(self.dll.b2f_stream_write32(C.c_void_p(dptr), value, C.c_void_p(stream)))

    def stream_wait_geq32(self, dptr: int, value: int, stream: int = 0):
        self.check(self.dll.b2f_stream_wait_geq32(C.c_void_p(dptr), value, C.c_void_p(stream)))

    def memset_async(self, dptr: int, byte: int, nbytes: int, stream: int = 0):
        self.check(self.dll.b2f_memset_async(C.c_void_p(dptr), byte, nbytes, C.c_void_p(stream)))

    # -- local extrema / blob_LoG plumbing -----------------------------------------------------
    def findlocalextrema(self, img: b2f_array, minima: bool, window, edges, stream: int = 0) -> np.ndarray:
        """-> 0-based column-major linear indices of the peaks, ascending (int64)."""
        nd = img.ndim
        win = (C.c_int64 * MAXDIM)(*list(window) + [1] * (MAXDIM - nd))
        edg = (C.c_int32 * MAXDIM)(*[1 if e else 0 for e in edges] + [1] * (MAXDIM - nd))
        n = 1
        for d in range(nd):
            n *= img.dims[d]
        cap = max(1024, n // 8)         # a retry recomputes the scan: start above any realistic peak density
        while True:
            idx = np.empty(cap, dtype=np.int64)
            cnt = C.c_int64()
            self.check(self.dll.b2f_findlocalextrema(C.byref(img), 1 if minima else 0, win, edg,
                                                     idx.ctypes.data_as(C.POINTER(C.c_int64)), cap, C.byref(cnt),
                                                     C.c_void_p(stream)))
            if cnt.value <= cap:
                return idx[:cnt.value].copy()
            cap = int(cnt.value)

    def scale_into_slice(self, src: b2f_array, stack: b2f_array, slice_: int, scale: float, stream: int = 0):
        self.check(self.dll.b2f_scale_into_slice(C.byref(src), C.byref(stack), slice_, scale, C.c_void_p(stream)))

    def maxabs(self, img: b2f_array, stream: int = 0) -> float:
        r = C.c_double()
        self.check(self.dll.b2f_maxabs(C.byref(img), C.byref(r), C.c_void_p(stream)))
        return float(r.value)

    def gather(self, arr: b2f_array, idx: np.ndarray, stream: int = 0) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        out = np.empty(idx.size, dtype=np.float64)
        self.check(self.dll.b2f_gather(C.byref(arr), idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size,
                                       out.ctypes.data_as(C.POINTER(C.c_double)), C.c_void_p(stream)))
        return out

    # -- NA() border pieces --------------------------------------------------------------------
    def na_prepare(self, img: b2f_array, na_mode: int, imgtmp=None, valid=None, stream: int = 0) -> bool:
        h = C.c_int32()
        self.check(self.dll.b2f_na_prepare(C.byref(img), na_mode, C.byref(imgtmp) if imgtmp is not None else None,
                                           C.byref(valid) if valid is not None else None, C.byref(h), C.c_void_p(stream)))
        return bool(h.value)

    def divide(self, out: b2f_array, den: b2f_array, stream: int = 0):
        self.check(self.dll.b2f_divide(C.byref(out), C.byref(den), C.c_void_p(stream)))

    def normalize_dims(self, out: b2f_array, factors, stream: int = 0):
        keep = [np.ascontiguousarray(f, dtype=np.float64) for f in factors]
        arr = (C.POINTER(C.c_double) * len(keep))(*[k.ctypes.data_as(C.POINTER(C.c_double)) for k in keep])
        self.check(self.dll.b2f_normalize_dims(C.byref(out), arr, C.c_void_p(stream)))

    # -- raw device memory (arrays that never leave the GPU between calls) ------------------------
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self.check(self.dll.b2f_malloc(C.byref(p), max(1, int(nbytes))))
        return int(p.value)

    def free(self, dptr: int):
        self.check(self.dll.b2f_free(C.c_void_p(dptr)))
