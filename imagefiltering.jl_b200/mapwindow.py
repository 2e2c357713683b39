"""mapwindow / mapwindow! for f in {extrema, minimum, maximum} (reference src/mapwindow.jl:75-121,
337-481) and f = median! (the generic window loop :270-333 with Statistics.median!).  Arbitrary window functions are Julia closures and cannot cross a C ABI: they raise
NotSupportedError (there is no CPU fallback)."""
from __future__ import annotations

import builtins

import numpy as np

from . import _abi
from ._abi import ArgumentError, DimensionMismatch, NotSupportedError
from .border import Fill, Inner, NoPad, Pad, borderinstance
from .device import DeviceArray
from .imfilter import _as_input, _as_output
from .offsetarrays import OffsetArray


def extrema(a):
    """Marker mirroring Base.extrema (also usable on arrays)."""
    a = np.asarray(a)
    return a.min(), a.max()


def minimum(a):
    return np.asarray(a).min()


def maximum(a):
    return np.asarray(a).max()


def median(a):
    """Marker mirroring Statistics.median / median! (mapwindow(median!, ...) is the reference's spelling)."""
    return float(np.median(np.asarray(a)))


median_ = median        # `median!`


def mean(a):
    """Marker mirroring Statistics.mean."""
    return float(np.mean(np.asarray(a)))


def sum_(a):
    """Marker mirroring Base.sum (`sum` is a Python builtin, which is accepted as well)."""
    return np.asarray(a).sum()


_MED = {median, np.median, "median", "median!"}
_MIN = {minimum, builtins.min, np.min, np.amin, np.minimum.reduce, "minimum", "min"}
_MAX = {maximum, builtins.max, np.max, np.amax, np.maximum.reduce, "maximum", "max"}
_EXT = {extrema, "extrema"}
_MEAN = {mean, np.mean, "mean"}
_SUM = {sum_, builtins.sum, np.sum, "sum"}
_REDUCE_OP = {"median": _abi.WIN_MEDIAN, "mean": _abi.WIN_MEAN, "sum": _abi.WIN_SUM, "min": _abi.WIN_MIN, "max": _abi.WIN_MAX}


def _kind(f):
    try:
        if f in _EXT:
            return "extrema"
        if f in _MIN:
            return "min"
        if f in _MAX:
            return "max"
        if f in _MED:
            return "median"
        if f in _MEAN:
            return "mean"
        if f in _SUM:
            return "sum"
    except TypeError:
        pass
    raise NotSupportedError(
        "mapwindow on the device supports f in {extrema, minimum, maximum, median, mean, sum}; arbitrary window functions "
        "cannot cross the C ABI and there is no CPU fallback")


def resolve_window(window, ndim, allow_even=False):
    """src/mapwindow.jl:136-150 (+ the positional extrema fast path :337-338, which accepts even widths
    and places the window at [i-(w>>1), i-(w>>1)+w-1])."""
    if isinstance(window, (int, np.integer)):
        window = (int(window),)
    if isinstance(window, range):
        window = (window,)
    window = tuple(window)
    if len(window) == 0:
        raise ArgumentError("empty window")
    lo, hi = [], []
    for w in window:
        if isinstance(w, range):
            lo.append(w.start)
            hi.append(w.stop - 1)
        elif isinstance(w, (tuple, list)):
            lo.append(int(w[0]))
            hi.append(int(w[1]))
        else:
            w = int(w)
            if w < 1:
                raise ArgumentError(f"window entries must be positive, got {w}")
            if w % 2 == 0 and not allow_even:
                raise ArgumentError(f"entries in window must be odd, got {window}")
            h = w >> 1
            lo.append(-h)
            hi.append(-h + w - 1)
    if len(lo) != ndim:
        raise DimensionMismatch(f"window has {len(lo)} entries, image has {ndim} dimensions")
    return lo, hi


def _default_indices_ok(indices, first, shape):
    if indices is None:
        return True
    want = tuple(range(f, f + n) for f, n in zip(first, shape))
    return tuple(indices) == want


def _reduce_out_dtype(kind, dt):
    """Result eltype of the window reduction (compute_output_eltype, src/mapwindow.jl:253-257, for the supported f)."""
    isf = dt in (_abi.F32, _abi.F64)
    if kind in ("median", "mean"):
        return _abi.F32 if dt == _abi.F32 else _abi.F64
    if kind == "sum":
        return dt if isf else _abi.I64       # Julia widens small integers to Int / UInt
    return dt


def _mapwindow_indices(L, kind, out_spec, desc, ndim, first, shape, wlo, whi, b, indices):
    """mapwindow(f, img, window; indices=ranges): the window function is evaluated at the listed image indices only
    (src/mapwindow.jl:123-131,156-183,270-306).  The result has plain 1-based axes of the ranges' lengths (`axes(r, 1)` of an
    ordinary range is OneTo(length(r)), test/mapwindow.jl:146-151)."""
    if isinstance(indices, range):
        indices = (indices,)
    indices = tuple(indices)
    if len(indices) != ndim or not all(isinstance(r, range) for r in indices):
        raise ArgumentError("indices= takes one range per image dimension")
    if any(r.step < 1 for r in indices):
        raise NotSupportedError("indices= with decreasing ranges is outside the accelerated path")
    if kind == "extrema":
        raise NotSupportedError("mapwindow(extrema, ...; indices=...) is outside the accelerated path: take minimum and maximum separately")
    counts = tuple(len(r) for r in indices)
    for r, f0, n in zip(indices, first, shape):
        if len(r) and (r[0] < f0 or r[-1] > f0 + n - 1):
            raise DimensionMismatch("indices= must lie inside the image axes")
    odt = _reduce_out_dtype(kind, desc.dtype)
    idx_first, idx_step = [r.start for r in indices], [r.step for r in indices]
    if out_spec is not None:
        odesc, okeep = _as_output(out_spec)
        if tuple(odesc.dims[d] for d in range(ndim)) != counts:
            raise DimensionMismatch("out must have one element per requested index")
        od = _abi.make_array(odesc.ptr, odesc.dtype, counts, (1,) * ndim, odesc.mem)
        L.mapwindow_reduce(desc, od, _REDUCE_OP[kind], wlo, whi, b.to_abi(ndim), idx_first, idx_step)
        return out_spec
    res = np.empty(counts, dtype=_abi.DTYPE_TO_NP[odt], order="F")
    L.mapwindow_reduce(desc, _abi.make_array(res.ctypes.data, odt, counts, (1,) * ndim, _abi.HOST), _REDUCE_OP[kind], wlo, whi,
                       b.to_abi(ndim), idx_first, idx_step)
    return res


def _mapwindow(f, out_spec, img, window, border, indices, library):
    from ._lib import lib
    L = library if library is not None else lib()
    kind = _kind(f)
    desc, ndim, first, shape, keep = _as_input(img)
    fast = kind == "extrema" and border is None and indices is None
    wlo, whi = resolve_window(window, ndim, allow_even=fast)
    b = borderinstance("replicate" if border is None else border)
    if isinstance(b, NoPad):
        raise NotSupportedError("NoPad() is not supported by mapwindow (marked broken in the reference tests)")
    if isinstance(b, Inner):
        lo = [f0 - l for f0, l in zip(first, wlo)]
        hi = [f0 + n - 1 - h for f0, n, h in zip(first, shape, whi)]
        if indices is not None and tuple(indices if not isinstance(indices, range) else (indices,)) != tuple(range(l, h + 1) for l, h in zip(lo, hi)):
            return _mapwindow_indices(L, kind, out_spec, desc, ndim, first, shape, wlo, whi, b, indices)
    else:
        if not _default_indices_ok(indices if not isinstance(indices, range) else (indices,), first, shape):
            return _mapwindow_indices(L, kind, out_spec, desc, ndim, first, shape, wlo, whi, b, indices)
        lo, hi = list(first), [f0 + n - 1 for f0, n in zip(first, shape)]
    oshape = tuple(max(0, h - l + 1) for l, h in zip(lo, hi))
    base = _abi.DTYPE_TO_NP[desc.dtype]
    if kind in ("median", "mean", "sum"):        # generic window path, src/mapwindow.jl:270-333 with f = median! / mean / sum
        odt = _reduce_out_dtype(kind, desc.dtype)
        if out_spec is not None:
            odesc, okeep = _as_output(out_spec)
            od = _abi.make_array(odesc.ptr, odesc.dtype, oshape, lo, odesc.mem)
            L.mapwindow_reduce(desc, od, _REDUCE_OP[kind], wlo, whi, b.to_abi(ndim))
            return out_spec
        res = np.empty(oshape, dtype=_abi.DTYPE_TO_NP[odt], order="F")
        L.mapwindow_reduce(desc, _abi.make_array(res.ctypes.data, odt, oshape, lo, _abi.HOST), _REDUCE_OP[kind], wlo, whi, b.to_abi(ndim))
        return OffsetArray.with_first(res, lo) if any(l != 1 for l in lo) else res

    if out_spec is not None:  # mapwindow!
        out = out_spec
        odesc, okeep = _as_output(out if not (isinstance(out, np.ndarray) and out.dtype.names) else out.view(base).reshape((2,) + out.shape, order="F"))
        if kind == "extrema":
            od = _abi.make_array(odesc.ptr, desc.dtype, oshape, lo, odesc.mem)
            L.mapwindow_extrema(desc, od, None, True, wlo, whi, b.to_abi(ndim))
        else:
            od = _abi.make_array(odesc.ptr, desc.dtype, oshape, lo, odesc.mem)
            L.mapwindow_extrema(desc, od if kind == "min" else None, od if kind == "max" else None, False,
                                wlo, whi, b.to_abi(ndim))
        return out

    if kind == "extrema":
        pair = np.dtype([("min", base), ("max", base)])
        res = np.empty(oshape, dtype=pair, order="F")
        od = _abi.make_array(res.ctypes.data, desc.dtype, oshape, lo, _abi.HOST)
        L.mapwindow_extrema(desc, od, None, True, wlo, whi, b.to_abi(ndim))
    else:
        res = np.empty(oshape, dtype=base, order="F")
        od = _abi.make_array(res.ctypes.data, desc.dtype, oshape, lo, _abi.HOST)
        L.mapwindow_extrema(desc, od if kind == "min" else None, od if kind == "max" else None, False,
                            wlo, whi, b.to_abi(ndim))
    if any(l != 1 for l in lo):
        return OffsetArray.with_first(res, lo)
    return res


def mapwindow(f, img, window, border=None, indices=None, *, _library=None):
    """mapwindow(f, img, window; border="replicate", indices=axes(img))  (src/mapwindow.jl:75-85).
    `mapwindow(extrema, A, window)` without keywords is the reference's extrema_filter fast path
    (:337-338): even window widths are allowed there.  extrema returns a structured array with
    fields "min" and "max" laid out like Julia's Array{Tuple{T,T}}."""
    return _mapwindow(f, None, img, window, border, indices, _library)


def mapwindow_(f, out, img, window, border=None, indices=None, *, _library=None):
    """mapwindow!(f, out, img, window; border, indices)  (src/mapwindow.jl:107-121)."""
    return _mapwindow(f, out, img, window, "replicate" if border is None else border, indices, _library)
