"""imfilter / imfilter! / imgradients — host-side mirror of the reference's dispatch ladder
(src/imfilter.jl:2-49 `imfilter`, :204-254 `imfilter!`, src/specialty.jl:39-53 `imgradients`).

Everything here is metadata work (steps 1-5 of the reference: output eltype, kernel
canonicalisation, default border, allocation, resource check).  All arithmetic happens behind the
C ABI of libb2f.so (include/b2f.h); there is no numpy/CPU computation of filter results here.

Julia's `imfilter!` is spelled `imfilter_` (a trailing `!` is not a Python identifier).
Array convention: numpy axis k == Julia dimension k+1; data is handed to the library in Julia
(column-major) memory order, so C-ordered inputs are copied to Fortran order first.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _abi
from ._abi import ArgumentError, DimensionMismatch, InexactError, NotSupportedError
from .border import AbstractBorder, Fill, Inner, NA, NoPad, Pad, borderinstance
from .color import ColorArray, lift_kernel
from .device import DeviceArray
from .kernel import Laplacian
from .kernelfactors import ReshapedOneD, TriggsSdika
from .n0f8 import N0f8Array, n0f8
from .offsetarrays import OffsetArray, centered
from .resources import AbstractResource, Alg, CUDALibs, FFT, FIR, FIRTiled, IIR

STAGE_LAPLACIAN = 2


# ---- step 1: output element type (src/imfilter.jl:1131-1154) -----------------------------------
def _img_eltype(img):
    if isinstance(img, N0f8Array):
        return "n0f8"
    if isinstance(img, DeviceArray):
        return "n0f8" if img.dtype == _abi.N0F8 else _abi.DTYPE_TO_NP[img.dtype]
    if isinstance(img, OffsetArray):
        return _img_eltype(img.parent)
    return np.asarray(img).dtype


def _mul_type(S, K):
    """typeof(zero(S)*zero(K) + zero(S)*zero(K)) for the element types this package handles."""
    K = np.dtype(K)
    if isinstance(S, str):  # N0f8
        if K.kind == "f":
            return K
        return np.dtype(np.float32)  # floattype(N0f8); integer taps on N0f8 data (unpinned corner)
    S = np.dtype(S)
    if S.kind == "b":
        S = np.dtype(np.int8)
    if K.kind == "b":
        K = np.dtype(np.int64)
    if S.kind == "f" and K.kind in "iu":
        return S
    if K.kind == "f" and S.kind in "iu":
        return K
    return np.result_type(S, K)


def _kernel_dtype(k):
    if isinstance(k, Laplacian):
        return None
    if isinstance(k, TriggsSdika):
        return k.dtype
    if isinstance(k, (ReshapedOneD, OffsetArray)):
        return k.dtype
    return np.asarray(k).dtype


def filter_type(img, kernel):
    S = img if isinstance(img, (np.dtype, type, str)) else _img_eltype(img)
    ks = kernel if isinstance(kernel, tuple) else (kernel,)
    T = None
    for k in ks:
        if isinstance(k, Laplacian):  # src/imfilter.jl:1134-1139
            if isinstance(S, str):
                t = np.dtype(np.float32)
            else:
                s = np.dtype(S)
                t = {"u1": np.dtype(np.int16), "u2": np.dtype(np.int32), "u4": np.dtype(np.int64)}.get(
                    s.str[1:], np.dtype(np.int8) if s.kind == "b" else s)
        else:
            t = _mul_type(S, _kernel_dtype(k))
        T = t if T is None else np.promote_types(T, t)
    return T if T is not None else (np.dtype(np.float32) if isinstance(S, str) else np.dtype(S))


# ---- step 2: kernel canonicalisation (src/imfilter.jl:1156-1196) ---------------------------------
def _kernelshift(k):
    if isinstance(k, OffsetArray):
        return k
    warnings.warn("assuming that the origin is at the center of the kernel; to avoid this warning, "
                  "call `centered(kernel)` or use an OffsetArray", DeprecationWarning, stacklevel=4)
    return centered(np.asarray(k))


def factorkernel(kernel, integer_path=False):
    if isinstance(kernel, (Laplacian, ReshapedOneD)):
        return (kernel,)
    ks = _kernelshift(kernel)
    if ks.ndim != 2 or integer_path or ks.dtype.kind in "iub":
        return (ks,)
    # factorstridedkernel: LAPACK svd, separable iff all but the first singular value < sqrt(eps)
    U, S, Vt = np.linalg.svd(ks.parent.astype(np.float64 if ks.dtype != np.float32 else np.float32))
    eps = np.sqrt(np.finfo(S.dtype).eps)
    if np.any(np.abs(S[1:]) >= eps):
        dummy = OffsetArray.with_first(np.ones((1, 1), dtype=np.int64), (0, 0))
        return (dummy, ks)
    ss = np.sqrt(S[0])
    u = (ss * U[:, :1]).astype(S.dtype)
    v = (ss * Vt[:1, :]).astype(S.dtype)
    return (OffsetArray.with_first(u, (ks.first[0], 0)), OffsetArray.with_first(v, (0, ks.first[1])))


def _tap_dtype(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        return _abi.TAPS_F32 if dt == np.float32 else _abi.TAPS_F64
    if dt.kind in "iub":
        return _abi.TAPS_INT
    raise NotSupportedError(f"kernel eltype {dt} is not supported")


def build_stages(kernel, ndim):
    """ProcessedKernel tuple -> list of stage dicts for _abi.StageList."""
    stages = []
    for k in kernel:
        if isinstance(k, Laplacian):
            if k.ndim != ndim:
                raise DimensionMismatch(f"Laplacian has {k.ndim} dims, image has {ndim}")
            stages.append(dict(kind=STAGE_LAPLACIAN, ndim=ndim, tap_dtype=_abi.TAPS_INT,
                               len=[3 if f else 1 for f in k.flags], lo=[-1 if f else 0 for f in k.flags],
                               taps=np.zeros(1)))
            continue
        if isinstance(k, ReshapedOneD):
            if k.N != ndim:
                raise DimensionMismatch(f"kernel factor is for {k.N}-d arrays, image has {ndim} dims")
            ln = [1] * ndim
            lo = [0] * ndim
            ln[k.Npre] = k.data.shape[0]
            lo[k.Npre] = k.data.first[0]
            stages.append(dict(kind=_abi.STAGE_1D, axis=k.Npre, ndim=ndim, tap_dtype=_tap_dtype(k.dtype),
                               len=ln, lo=lo, taps=k.data.parent, reshaped=True))
            continue
        if isinstance(k, OffsetArray):
            p, first = k.parent, list(k.first)
        else:  # plain array inside a tuple: axes are 1:n (samedims -> reshape, no centring)
            p = np.asarray(k)
            first = [1] * p.ndim
        if p.ndim > ndim:
            raise DimensionMismatch(f"kernel has {p.ndim} dims, image has {ndim}")
        extra = ndim - p.ndim
        shape = list(p.shape) + [1] * extra
        first = first + [0 if isinstance(k, OffsetArray) else 1] * extra
        ext = [d for d in range(ndim) if shape[d] > 1]
        td = _tap_dtype(p.dtype)
        if len(ext) <= 1 and all(first[d] == 0 for d in range(ndim) if d not in ext):
            ax = ext[0] if ext else 0
            stages.append(dict(kind=_abi.STAGE_1D, axis=ax, ndim=ndim, tap_dtype=td, len=shape, lo=first,
                               taps=p.reshape(-1, order="F")))
        else:
            stages.append(dict(kind=_abi.STAGE_DENSE, ndim=ndim, tap_dtype=td, len=shape, lo=first,
                               taps=np.asarray(p).reshape(shape, order="F")))
    return stages


def kernel_extent(stages, ndim):
    """accumulate_padding (src/border.jl:614-642): summed first/last tap index per axis."""
    first = [0] * ndim
    last = [0] * ndim
    for s in stages:
        for d in range(ndim):
            if s["kind"] == _abi.STAGE_1D and d != s["axis"]:
                continue
            first[d] += s["lo"][d]
            last[d] += s["lo"][d] + s["len"][d] - 1
    return first, last


# ---- array plumbing -----------------------------------------------------------------------------
def _as_input(img):
    """-> (descriptor, ndim, first, shape, keepalive)"""
    if isinstance(img, DeviceArray):
        return img.desc(), img.ndim, img.origin, img.dims, img
    first = None
    if isinstance(img, OffsetArray):
        first, img = img.first, img.parent
    dt = None
    if isinstance(img, N0f8Array):
        img, dt = img.raw, _abi.N0F8
    a = np.asarray(img)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype not in _abi.NP_TO_DTYPE:
        raise NotSupportedError(f"image eltype {a.dtype} is not supported")
    a = np.asfortranarray(a) if a.ndim > 1 else np.ascontiguousarray(a)
    first = tuple(first) if first is not None else (1,) * a.ndim
    return _abi.numpy_array_desc(a, first, dt), a.ndim, first, a.shape, a


def _as_output(out):
    if isinstance(out, DeviceArray):
        return out.desc(), out
    first = None
    if isinstance(out, OffsetArray):
        first, out = out.first, out.parent
    if not isinstance(out, np.ndarray):
        raise TypeError("out must be a numpy array, OffsetArray or DeviceArray")
    if out.ndim > 1 and not out.flags.f_contiguous:
        raise ValueError("out must be Fortran-contiguous (Julia memory order)")
    return _abi.numpy_array_desc(out, first), out


def allocate_output(T, img_first, img_shape, stages, border):
    """src/border.jl:686-696."""
    ndim = len(img_shape)
    T = np.dtype(T)
    if isinstance(border, Inner):
        if border.lo:
            if len(border.lo) != ndim:
                raise DimensionMismatch(f"dimensionality of img and the border must agree, got {ndim} and {len(border.lo)}")
            lo = [f + l for f, l in zip(img_first, border.lo)]
            hi = [f + n - 1 - h for f, n, h in zip(img_first, img_shape, border.hi)]
        else:
            kf, kl = kernel_extent(stages, ndim)
            lo = [max(f, f - a) for f, a in zip(img_first, kf)]
            hi = [min(f + n - 1, f + n - 1 - b) for f, n, b in zip(img_first, img_shape, kl)]
        shape = tuple(max(0, h - l + 1) for l, h in zip(lo, hi))
        return OffsetArray.with_first(np.empty(shape, dtype=T, order="F"), lo)
    arr = np.empty(img_shape, dtype=T, order="F")
    if any(f != 1 for f in img_first):
        return OffsetArray.with_first(arr, img_first)
    return arr


# ---- argument ladder ------------------------------------------------------------------------------
def _is_type(x):
    return isinstance(x, (type, np.dtype)) or (isinstance(x, str) and x in ("float32", "float64"))


def _check_resource(r):
    if not isinstance(r, CUDALibs):
        raise NotSupportedError(
            f"{r!r}: this package accelerates CUDALibs(Algorithm.FIR()) only and has no CPU execution path")
    if not isinstance(r.settings, (FIR, FIRTiled)):
        raise NotSupportedError(f"{r!r}: only Algorithm.FIR() is implemented on the device (FFT/IIR are out of scope)")
    return r


def _split_tail(args):
    border, alg = "replicate", None
    rest = list(args)
    if rest and isinstance(rest[0], (str, AbstractBorder)):
        border = rest.pop(0)
    if rest and isinstance(rest[0], Alg):
        alg = rest.pop(0)
    if rest:
        raise TypeError(f"no method matching imfilter(…, {rest})")
    return borderinstance(border), alg


# ---- IIR (Triggs-Sdika) kernels: src/imfilter.jl:890-1092 -------------------------------------------------------------
def _iir_factors(kernel, ndim=None):
    """-> [(axis, TriggsSdika), ...] when every factor of `kernel` is an IIR filter, None when none is (the FIR ladder goes
    on).  A bare TriggsSdika in position d of a tuple filters axis d (_imfilter_inplace_tuple!, :946-960); a ReshapedOneD
    carries its axis.  Tuples mixing FIR and IIR factors are the reference's Algorithm.Mixed (outside this package)."""
    ks = kernel if isinstance(kernel, tuple) else (kernel,)
    is_iir = [isinstance(k, TriggsSdika) or (isinstance(k, ReshapedOneD) and isinstance(k.data, TriggsSdika)) for k in ks]
    if not any(is_iir):
        return None
    if not all(is_iir):
        raise NotSupportedError("kernels mixing FIR and IIR factors (Algorithm.Mixed) are outside the accelerated path")
    out = []
    for d, k in enumerate(ks):
        if isinstance(k, ReshapedOneD):
            if ndim is not None and k.N != ndim:
                raise DimensionMismatch(f"kernel factor is for {k.N}-d arrays, image has {ndim} dims")
            out.append((k.Npre, k.data))
        else:
            out.append((d, k))
    if ndim is not None and len(out) > ndim:
        raise DimensionMismatch("cannot have more kernels than dimensions")
    return out


def _iir_border(border):
    if isinstance(border, NA):
        return border
    if isinstance(border, Fill):
        return border
    if isinstance(border, Pad) and border.style == "replicate":
        return border
    raise ArgumentError('only "replicate" is supported')                 # src/imfilter.jl:897


def _iir_abi_border(border, ndim):
    return (Fill(border.value) if isinstance(border, Fill) else Pad("replicate")).to_abi(ndim)


def _run_iir(L, odesc, img_desc, ndim, factors, border):
    """The cascade: the first factor reads img, the later ones filter out in place (:946-960)."""
    if isinstance(border, NA):
        return _run_iir_na(L, odesc, img_desc, ndim, factors, border)
    b = _iir_abi_border(border, ndim)
    src = img_desc
    if not factors:
        raise ArgumentError("empty IIR kernel")
    for axis, k in factors:
        L.iir(src, odesc, axis, k.coefficients(), b)
        src = odesc


def _run_iir_na(L, odesc, img_desc, ndim, factors, border):
    """imfilter!(r, out, img, kernel::Tuple{IIR...}, NA(na)): src/imfilter.jl:282-318 with the IIR methods :1094-1108 (NaNs:
    zero them, filter image and validity mask in place with Fill(0), divide) and :1222-1232 (no NaNs: filter with Fill(0),
    divide by the per-axis responses to a vector of ones)."""
    if odesc.dtype not in (_abi.F32, _abi.F64):
        raise NotSupportedError("NA() needs a Float32 / Float64 output")
    dims = [int(odesc.dims[d]) for d in range(ndim)]
    can_na = img_desc.dtype in (_abi.F32, _abi.F64)
    hasna = L.na_prepare(img_desc, border.mode) if can_na else False
    fill0 = Fill(0)
    if not hasna:
        if len(factors) != ndim:
            raise TypeError("MethodError: no method matching normalize_separable! (one kernel factor per dimension)")
        _run_iir(L, odesc, img_desc, ndim, factors, fill0)
        facs = [None] * ndim
        for axis, k in factors:
            ones = np.ones(dims[axis])
            d1 = _abi.numpy_array_desc(ones, (1,))
            L.iir(d1, d1, 0, k.coefficients(), fill0.to_abi(1))
            facs[axis] = ones
        L.normalize_dims(odesc, facs)
        return
    n = int(np.prod(dims))
    origin = [int(img_desc.origin[d]) for d in range(ndim)]
    if img_desc.mem == _abi.DEVICE:
        p = L.malloc(n * _abi.DTYPE_SIZE[odesc.dtype])
        valid = _abi.make_array(p, odesc.dtype, dims, origin, _abi.DEVICE)
    else:
        p = None
        va = np.empty(dims, dtype=_abi.DTYPE_TO_NP[odesc.dtype], order="F")
        valid = _abi.numpy_array_desc(va, origin)
    try:
        L.na_prepare(img_desc, border.mode, odesc, valid)
        _run_iir(L, odesc, odesc, ndim, factors, fill0)
        _run_iir(L, valid, valid, ndim, factors, fill0)
        L.divide(odesc, valid)
    finally:
        if p is not None:
            L.check(L.dll.b2f_sync())
            L.free(p)


# ---- FFT algorithm: src/imfilter.jl:776-888 ---------------------------------------------------------------------------------
def kernelconv(kernel, ndim):
    """kernelconv(kernel...) (src/imfilter.jl:1257-1280): the one dense kernel a cascade of factors is equivalent to — the full
    convolution of the factors, first indices adding up.  Kernel construction on the host, like factorkernel's SVD."""
    from scipy.signal import convolve
    acc, first = None, None
    for k in kernel:
        if isinstance(k, (Laplacian, TriggsSdika)) or (isinstance(k, ReshapedOneD) and isinstance(k.data, TriggsSdika)):
            raise NotSupportedError(f"{type(k).__name__} kernels have no array form for the FFT algorithm here")
        if isinstance(k, ReshapedOneD):
            if k.N != ndim:
                raise DimensionMismatch(f"kernel factor is for {k.N}-d arrays, image has {ndim} dims")
            d = k.dense()
            p, f = d.parent, tuple(d.first)
        elif isinstance(k, OffsetArray):
            p, f = k.parent, tuple(k.first)
            if p.ndim > ndim:
                raise DimensionMismatch(f"kernel has {p.ndim} dims, image has {ndim}")
            p, f = p.reshape(p.shape + (1,) * (ndim - p.ndim)), f + (0,) * (ndim - p.ndim)
        else:                                   # a plain array inside a tuple keeps its axes 1:n
            p = np.asarray(k)
            if p.ndim > ndim:
                raise DimensionMismatch(f"kernel has {p.ndim} dims, image has {ndim}")
            f = (1,) * ndim
            p = p.reshape(p.shape + (1,) * (ndim - p.ndim))
        if acc is None:
            acc, first = np.array(p), f
        else:
            acc = convolve(acc, p, mode="full", method="direct")
            first = tuple(a + b for a, b in zip(first, f))
    return OffsetArray.with_first(np.asfortranarray(acc), first)


def _is_fft(r, alg):
    return isinstance(alg, FFT) or (r is not None and isinstance(getattr(r, "settings", None), FFT))


def _run_fft(L, out, desc, ndim, kernel, border, roi):
    if isinstance(border, NA):
        raise NotSupportedError("NA() with Algorithm.FFT() is not available")
    kc = kernelconv(kernel, ndim)
    p = kc.parent
    stage = dict(kind=_abi.STAGE_DENSE, ndim=ndim, tap_dtype=_tap_dtype(p.dtype), len=list(p.shape), lo=list(kc.first),
                 taps=np.asarray(p, dtype=np.float64).reshape(p.shape, order="F"))
    odesc, keep = _as_output(out)
    if odesc.ndim != ndim:
        raise DimensionMismatch(f"out has {odesc.ndim} dims, img has {ndim}")
    L.imfilter_fft(desc, odesc, _abi.StageList([stage]), border.to_abi(ndim), roi)
    return stage


def imfilter(*args, _library=None):
    """imfilter([r], [T], img, kernel, [border], [alg]) -> filtered array   (src/imfilter.jl:2-49)."""
    args = list(args)
    r = args.pop(0) if args and isinstance(args[0], AbstractResource) else None
    T = args.pop(0) if args and _is_type(args[0]) else None
    if len(args) < 2:
        raise TypeError("imfilter needs at least (img, kernel)")
    img, kernel = args[0], args[1]
    border, alg = _split_tail(args[2:])
    if r is not None and alg is not None:
        raise TypeError("MethodError: a resource and an algorithm cannot both be given")
    r_given = r is not None
    iir = _iir_factors(kernel)
    if iir is None and _is_fft(r, alg) and not isinstance(img, ColorArray):
        if r is not None and not isinstance(r, CUDALibs):
            raise NotSupportedError(f"{r!r}: this package has no CPU execution path")
        ks = kernel if isinstance(kernel, tuple) else (_kernelshift(kernel) if not isinstance(kernel, (Laplacian, ReshapedOneD)) else kernel,)
        T = np.dtype(T if T is not None else filter_type(img, ks))
        if T.kind != "f":
            raise InexactError(f"Algorithm.FFT() produces floating-point values; eltype {T} cannot hold them (ask for Float64)")
        desc, ndim, first, shape, keep = _as_input(img)
        st = [dict(kind=_abi.STAGE_DENSE, ndim=ndim, len=list(kernelconv(ks, ndim).parent.shape), lo=list(kernelconv(ks, ndim).first))]
        out = allocate_output(T, first, shape, st, border)
        from ._lib import lib
        _run_fft(_library if _library is not None else lib(), out, desc, ndim, ks, border, None)
        return out
    r = _resolve_resource(r, alg, iir is not None)
    if isinstance(img, ColorArray) and iir is not None:
        # colour image, IIR kernel (test/triggs.jl:45-60, imgc): every channel is filtered along the spatial axes, i.e. the
        # channel-leading (C, dims...) array along axes 1 .. N; Fill(value) fills every channel with the same value
        raw = n0f8(img.data) if img.data.dtype == np.uint8 else img.data
        T = np.dtype(T if T is not None else filter_type(raw, kernel))
        desc, ndim, first, shape, keep = _as_input(raw)
        factors = [(ax + 1, k) for ax, k in _iir_factors(kernel, ndim - 1)]
        out = allocate_output(T, first, shape, [], Pad("replicate"))
        from ._lib import lib
        b = _iir_border(border)
        if isinstance(b, NA):
            raise NotSupportedError("NA() on colour images with IIR kernels is not available")
        _run_iir(_library if _library is not None else lib(), _as_output(out)[0], desc, ndim, factors, b)
        return ColorArray(out)
    if isinstance(img, ColorArray):
        return _imfilter_color(r, T, img, kernel, border, _library)
    if iir is not None:
        T = np.dtype(T if T is not None else filter_type(img, kernel))
        desc, ndim, first, shape, keep = _as_input(img)
        factors = _iir_factors(kernel, ndim)
        out = allocate_output(T, first, shape, [], Pad("replicate"))
        from ._lib import lib
        _run_iir(_library if _library is not None else lib(), _as_output(out)[0], desc, ndim, factors, _iir_border(border))
        return out
    if T is None:
        T = filter_type(img, kernel)
    T = np.dtype(T)
    S = _img_eltype(img)
    if not isinstance(kernel, tuple):
        kd = _kernel_dtype(kernel)
        int_path = (T.kind in "iub" and not isinstance(S, str) and np.dtype(S).kind in "iub"
                    and kd is not None and np.dtype(kd).kind in "iub")
        if int_path and not r_given and not isinstance(kernel, Laplacian):
            # src/imfilter.jl:10-12: T, TI, TK <: Integer (no resource given) wraps the kernel as `(kernel,)` WITHOUT
            # kernelshift — a plain array keeps its axes 1:n (out[i] = sum_j A[i+j] k[j], j = 1..n), an OffsetArray its own
            kernel = (kernel,)
        else:
            kernel = factorkernel(kernel, int_path)
    desc, ndim, first, shape, keep = _as_input(img)
    stages = build_stages(kernel, ndim)
    out = allocate_output(T, first, shape, stages, border)
    _run(r, out, desc, ndim, stages, border, None, _library)
    return out


def _imfilter_color(r, T, img, kernel, border, library):
    """imfilter on a colour image (RGB{N0f8}, RGB{Float32}, ...): every channel with the same kernel (the reference's
    eltype arithmetic, src/imfilter.jl:1131-1154) = the N+1-d path with the kernel lifted past the channel axis."""
    raw = n0f8(img.data) if img.data.dtype == np.uint8 else img.data
    if not isinstance(kernel, tuple):
        kernel = factorkernel(kernel)
    lifted = lift_kernel(kernel)
    if T is None:
        T = filter_type(raw, lifted)            # RGB{N0f8} * Float64 -> RGB{Float64}: the channel type follows the scalar rule
    T = np.dtype(T)
    if isinstance(border, (Pad, Fill, Inner)) and border.lo:
        border = type(border)(*((border.value,) if isinstance(border, Fill) else (border.style,) if isinstance(border, Pad) else ()),
                              (0,) + tuple(border.lo), (0,) + tuple(border.hi))
    desc, ndim, first, shape, keep = _as_input(raw)
    stages = build_stages(lifted, ndim)
    out = allocate_output(T, first, shape, stages, border)
    _run(r, out, desc, ndim, stages, border, None, library, nlead=1)
    if isinstance(out, OffsetArray):            # Inner(): the spatial axes shrink; the channel axis keeps index 1
        return out
    return ColorArray(out)


def _resolve_resource(r, alg, iir=False):
    """filter_algorithm (src/imfilter.jl:1200-1209): all-IIR kernels run Algorithm.IIR(), everything else FIR."""
    if iir:
        given = r.settings if r is not None else alg
        if r is not None and not isinstance(r, CUDALibs):
            raise NotSupportedError(f"{r!r}: this package has no CPU execution path")
        if given is not None and not isinstance(given, IIR):
            raise TypeError(f"MethodError: no method matching imfilter! for an IIR kernel with {given!r}")
        return CUDALibs(IIR())
    if r is None:
        if alg is not None and not isinstance(alg, (FIR, FIRTiled)):
            raise NotSupportedError(f"{alg!r} is outside the accelerated FIR path")
        return CUDALibs(FIR())
    return _check_resource(r)


def imfilter_(*args, _library=None):
    """imfilter!([r], out, img, kernel, [border], [alg | inds])   (src/imfilter.jl:204-254, 367-395)."""
    args = list(args)
    r = args.pop(0) if args and isinstance(args[0], AbstractResource) else None
    if len(args) < 3:
        raise TypeError("imfilter_ needs at least (out, img, kernel)")
    out, img, kernel = args[0], args[1], args[2]
    rest = args[3:]
    inds = None
    if rest and isinstance(rest[-1], (tuple, list)) and rest[-1] and isinstance(rest[-1][0], (range, tuple, list)):
        inds = rest.pop()
    dim = None
    if isinstance(kernel, TriggsSdika) and rest and isinstance(rest[0], (int, np.integer)):
        dim = int(rest.pop(0))                  # imfilter!(r, out, img, kernel::TriggsSdika, dim, border), 1-based (:922)
    border, alg = _split_tail(rest)
    if r is not None and alg is not None:
        raise TypeError("MethodError: a resource and an algorithm cannot both be given")
    iir = _iir_factors(kernel)
    if iir is None and _is_fft(r, alg):
        if r is not None and not isinstance(r, CUDALibs):
            raise NotSupportedError(f"{r!r}: this package has no CPU execution path")
        ks = kernel if isinstance(kernel, tuple) else (_kernelshift(kernel) if not isinstance(kernel, (Laplacian, ReshapedOneD)) else kernel,)
        desc, ndim, first, shape, keep = _as_input(img)
        roi = None
        if inds is not None:
            roi = ([(i.start if isinstance(i, range) else i[0]) for i in inds], [(i.stop - 1 if isinstance(i, range) else i[1]) for i in inds])
        from ._lib import lib
        _run_fft(_library if _library is not None else lib(), out, desc, ndim, ks, border, roi)
        return out
    r = _resolve_resource(r, alg, iir is not None)
    if iir is not None:
        desc, ndim, first, shape, keep = _as_input(img)
        factors = [(dim - 1, kernel)] if dim is not None else _iir_factors(kernel, ndim)
        if dim is not None and not 1 <= dim <= ndim:
            raise DimensionMismatch(f"dimension {dim} outside a {ndim}-d array")
        odesc, okeep = _as_output(out)
        if odesc.ndim != ndim or any(int(odesc.dims[d]) != int(desc.dims[d]) for d in range(ndim)):
            raise DimensionMismatch("out must have the axes of img")
        from ._lib import lib
        _run_iir(_library if _library is not None else lib(), odesc, desc, ndim, factors, _iir_border(border))
        return out
    if not isinstance(kernel, tuple):
        kernel = factorkernel(kernel)
    desc, ndim, first, shape, keep = _as_input(img)
    stages = build_stages(kernel, ndim)
    roi = None
    if inds is not None:
        lo = [(i.start if isinstance(i, range) else i[0]) for i in inds]
        hi = [(i.stop - 1 if isinstance(i, range) else i[1]) for i in inds]
        roi = (lo, hi)
    _run(r, out, desc, ndim, stages, border, roi, _library)
    return out


def _run(r, out, img_desc, ndim, stages, border, roi, library, nlead=0):
    from ._lib import lib
    L = library if library is not None else lib()
    odesc, keep = _as_output(out)
    if odesc.ndim != ndim:
        raise DimensionMismatch(f"out has {odesc.ndim} dims, img has {ndim}")
    if isinstance(border, NA):
        if roi is not None:
            raise NotSupportedError("NA() with inds is not available")
        return _run_na(L, odesc, img_desc, ndim, stages, border, nlead)
    sl = _abi.StageList(stages)
    L.imfilter(img_desc, odesc, sl, border.to_abi(ndim), roi)


def _run_na(L, odesc, img_desc, ndim, stages, border, nlead=0):
    """imfilter!(r, out, img, kernel, NA(na))  (src/imfilter.jl:282-318): flags -> separable or inseparable NA filtering.
    Every array operation is a call into the library; this function only sequences them."""
    if odesc.dtype not in (_abi.F32, _abi.F64):
        raise NotSupportedError("NA() needs a Float32 / Float64 output (the renormalising division)")
    dims = [int(odesc.dims[d]) for d in range(ndim)]
    if [int(img_desc.dims[d]) for d in range(ndim)] != dims:
        raise DimensionMismatch("NA(): out must have the axes of img")
    sl = _abi.StageList(stages)
    fill0 = Fill(0).to_abi(ndim)
    can_na = img_desc.dtype in (_abi.F32, _abi.F64)                 # Integer / fixed-point images cannot hold NaN
    separable = all(st["kind"] == _abi.STAGE_1D and (st.get("reshaped") or sum(1 for n in st["len"] if n > 1) == 1)
                    for st in stages)                               # isseparable, src/imfilter.jl:1219
    hasna = L.na_prepare(img_desc, border.mode) if can_na else False
    if separable and not hasna:                                     # imfilter_na_separable!, :1123-1127
        if len(stages) != ndim - nlead:
            raise TypeError("MethodError: no method matching normalize_separable! (one kernel factor per dimension)")
        L.imfilter(img_desc, odesc, sl, fill0)
        factors = [np.ones(dims[d]) for d in range(nlead)]          # colour channels: nothing to normalise along them
        for d in range(nlead, ndim):                                # normalize_separable!, :1234-1239
            st = stages[d - nlead]
            ax = st["axis"]
            ones, res = np.ones(dims[d]), np.empty(dims[d])
            one_d = dict(kind=_abi.STAGE_1D, axis=0, ndim=1, tap_dtype=st["tap_dtype"], len=[st["len"][ax]],
                         lo=[st["lo"][ax]], taps=st["taps"])
            L.imfilter(_abi.numpy_array_desc(ones, (1,)), _abi.numpy_array_desc(res, (1,)), _abi.StageList([one_d]),
                       Fill(0).to_abi(1))
            factors.append(res)
        L.normalize_dims(odesc, factors)
        return
    # imfilter_na_inseparable!, :1110-1121: temporaries live where img lives
    n = int(np.prod(dims))
    esz = _abi.DTYPE_SIZE[odesc.dtype]
    origin = [int(img_desc.origin[d]) for d in range(ndim)]
    owned, keep = [], []
    try:
        def temp():
            if img_desc.mem == _abi.DEVICE:
                p = L.malloc(n * esz)
                owned.append(p)
                return _abi.make_array(p, odesc.dtype, dims, origin, _abi.DEVICE)
            a = np.empty(dims, dtype=_abi.DTYPE_TO_NP[odesc.dtype], order="F")
            keep.append(a)
            return _abi.numpy_array_desc(a, origin)
        imgtmp, valid, vp = temp(), temp(), temp()
        L.na_prepare(img_desc, border.mode if can_na else NA.MODES["never"], imgtmp, valid)
        L.imfilter(imgtmp, odesc, sl, fill0)
        L.imfilter(valid, vp, sl, fill0)
        L.divide(odesc, vp)
    finally:
        if owned:
            L.check(L.dll.b2f_sync())
        for p in owned:
            L.free(p)


def imgradients(img, kernelfun, border="replicate", *, _library=None, T=None):
    """imgradients(img, kernelfun, border) -> tuple of N gradient arrays  (src/specialty.jl:39-53).
    One fused device launch reads `img` once for all N planes."""
    from ._lib import lib
    L = _library if _library is not None else lib()
    border = borderinstance(border)
    desc, ndim, first, shape, keep = _as_input(img)
    extended = tuple(n > 1 for n in shape)
    planes = []
    all_stages = []
    nst = None
    for d in range(1, ndim + 1):
        kern = kernelfun(extended, d)
        if not isinstance(kern, tuple):
            kern = factorkernel(kern)
        st = build_stages(kern, ndim)
        if nst is None:
            nst = len(st)
        elif nst != len(st):
            raise NotSupportedError("gradient kernels with differing stage counts")
        Td = np.dtype(T) if T is not None else filter_type(img, kern)
        planes.append(allocate_output(Td, first, shape, st, border))
        all_stages += st
    odescs = [_as_output(p)[0] for p in planes]
    L.imgradients(desc, odescs, _abi.StageList(all_stages), nst, border.to_abi(ndim))
    return tuple(planes)


def padarray(*args, _library=None):
    """padarray([T], img, border) (src/border.jl:324-352) — on this package the padded copy is
    produced by the device library (a cascade of zero stages evaluated over the padded axes)."""
    args = list(args)
    T = args.pop(0) if _is_type(args[0]) else None
    img, border = args
    border = borderinstance(border)
    if not isinstance(border, (Pad, Fill)) or not border.lo:
        raise ArgumentError(f"{border!r} lacks the proper padding sizes for an array with {np.ndim(img) if not hasattr(img, 'ndim') else img.ndim} dimensions")
    desc, ndim, first, shape, keep = _as_input(img)
    if len(border.lo) != ndim:
        raise ArgumentError(f"{border!r} lacks the proper padding sizes for an array with {ndim} dimensions")
    S = _img_eltype(img)
    T = np.dtype(T) if T is not None else (np.dtype(np.float32) if isinstance(S, str) else np.dtype(S))
    lo = [f - l for f, l in zip(first, border.lo)]
    oshape = tuple(n + l + h for n, l, h in zip(shape, border.lo, border.hi))
    out = OffsetArray.with_first(np.empty(oshape, dtype=T, order="F"), lo)
    _run(None, out, desc, ndim, [], border, None, _library)
    return out



class accum_mode:
    """Context manager over b2f_set_accum_mode (include/b2f.h): `with accum_mode("fma"): imfilter(...)` lets the library fuse the
    multiply and the add of Float64 accumulations (SURVEY Appendix C typing: the reference's Float64 results come from separate
    multiplies and adds, which "exact" — the default — reproduces bit for bit)."""

    def __init__(self, mode, _library=None):
        self.mode = {"exact": _abi.ACCUM_EXACT, "fma": _abi.ACCUM_FMA}[mode] if isinstance(mode, str) else int(mode)
        self._library = _library

    def __enter__(self):
        from ._lib import lib
        self._lib = self._library if self._library is not None else lib()
        self.prev = self._lib.set_accum_mode(self.mode)
        return self

    def __exit__(self, *exc):
        self._lib.set_accum_mode(self.prev)
        return False
