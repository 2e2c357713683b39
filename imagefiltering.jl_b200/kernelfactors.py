"""KernelFactors — separable kernel constructors (reference src/kernelfactors.jl).

Host-side data producers: they build the taps, the device library only ever sees numbers.
A `ReshapedOneD(N, Npre, data)` is "1-D taps acting on axis Npre (0-based) of an N-d array"
(src/kernelfactors.jl:52-131).
"""
from __future__ import annotations

import math

import numpy as np

from ._abi import ArgumentError
from .offsetarrays import OffsetArray, centered


class TriggsSdika:
    """TriggsSdika(a, b, scale, M) / TriggsSdika(ab, scale)  (src/kernelfactors.jl:463-505): a recursive (IIR) 1-D filter with a
    forward filter `a`, a backward filter `b` (3 coefficients each), the Triggs-Sdika boundary matrix `M` (3 x 3) and a final
    `scale`.  Coefficients are numpy scalars of one float type T; every derived quantity is computed in T, in the reference's
    order of operations."""
    __slots__ = ("a", "b", "scale", "M", "asum", "bsum", "dtype")

    def __init__(self, a, b=None, scale=None, M=None, dtype=None):
        if scale is None:                            # TriggsSdika(ab, scale)
            b, scale = None, b
        T = np.dtype(dtype if dtype is not None else np.result_type(*[np.asarray(x).dtype for x in a])).type
        if np.dtype(T).kind != "f":
            T = np.float64
        a = tuple(T(x) for x in a)
        if len(a) != 3:
            raise ArgumentError("only length 3 filters are currently supported")
        if b is None:
            a1, a2, a3 = a
            one = T(1)
            Mdenom = (one + a1 - a2 + a3) * (one - a1 - a2 - a3) * (one + a2 + (a1 - a3) * a3)
            M = [[-a3 * a1 + one - a3 * a3 - a2, (a3 + a1) * (a2 + a3 * a1), a3 * (a1 + a3 * a2)],
                 [a1 + a3 * a2, -(a2 - one) * (a2 + a3 * a1), -(a3 * a1 + a3 * a3 + a2 - one) * a3],
                 [a3 * a1 + a2 + a1 * a1 - a2 * a2, a1 * a2 + a3 * (a2 * a2) - a1 * (a3 * a3) - a3 * a3 * a3 - a3 * a2 + a3,
                  a3 * (a1 + a3 * a2)]]
            M = [[T(x / Mdenom) for x in row] for row in M]
            b = a
        else:
            b = tuple(T(x) for x in b)
            if len(b) != 3:
                raise ArgumentError("only length 3 filters are currently supported")
            M = [[T(x) for x in row] for row in np.asarray(M).reshape(3, 3)]
        self.a, self.b, self.scale, self.M = a, b, T(scale), M
        self.asum = (a[0] + a[1]) + a[2]
        self.bsum = (b[0] + b[1]) + b[2]
        self.dtype = np.dtype(T)

    @property
    def ndim(self):
        return 1

    def iscopy(self):
        """src/imfilter.jl:1254"""
        return all(x == 0 for x in self.a) and all(x == 0 for x in self.b) and self.scale == 1

    def coefficients(self):
        """The 18 doubles of b2f_iir (include/b2f.h): a[3], b[3], scale, M[9] row-major, 1 - asum, 1 - bsum (both formed in T)."""
        T = self.dtype.type
        return np.array(list(self.a) + list(self.b) + [self.scale] + [x for row in self.M for x in row] +
                        [T(1) - self.asum, T(1) - self.bsum], dtype=np.float64)

    def __repr__(self):
        return f"TriggsSdika{{{self.dtype}}}(a={tuple(float(x) for x in self.a)}, scale={float(self.scale)})"


class ReshapedOneD:
    __slots__ = ("N", "Npre", "data")

    def __init__(self, N, Npre, data):
        if isinstance(data, TriggsSdika):            # an IIR factor acting on axis Npre (src/kernelfactors.jl:560-565, iirg)
            self.N, self.Npre, self.data = int(N), int(Npre), data
            return
        if not isinstance(data, OffsetArray):
            data = OffsetArray.with_first(np.asarray(data), (1,))
        if data.ndim != 1:
            raise ArgumentError("ReshapedOneD needs a vector")
        self.N, self.Npre, self.data = int(N), int(Npre), data

    @property
    def axis(self):
        return self.Npre

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def axes(self):
        ax = [range(0, 1)] * self.N
        if not isinstance(self.data, TriggsSdika):   # Base.axes1(::TriggsSdika) = 0:0
            ax[self.Npre] = self.data.axes[0]
        return tuple(ax)

    def dense(self):
        """The equivalent N-d OffsetArray (used by Kernel.* products)."""
        shape = [1] * self.N
        shape[self.Npre] = self.data.shape[0]
        first = [0] * self.N
        first[self.Npre] = self.data.first[0]
        return OffsetArray.with_first(self.data.parent.reshape(shape), first)

    def __repr__(self):
        return f"ReshapedOneD(N={self.N}, axis={self.Npre}, first={self.data.first[0]}, {self.data.parent!r})"


def kernelfactors(factors):
    """src/kernelfactors.jl:586-596: factor d of an N-tuple of vectors acts on axis d."""
    factors = tuple(factors)
    N = len(factors)
    if all((f.ndim if hasattr(f, "ndim") else np.asarray(f).ndim) == 1 for f in factors):
        return tuple(ReshapedOneD(N, d, f if isinstance(f, OffsetArray) else np.asarray(f))
                     for d, f in enumerate(factors))
    out = []
    for f in factors:  # general arrays: extend to N dims with trailing singleton axes
        p, first = (f.parent, f.first) if isinstance(f, OffsetArray) else (np.asarray(f), (1,) * np.asarray(f).ndim)
        extra = N - p.ndim
        out.append(OffsetArray.with_first(p.reshape(p.shape + (1,) * extra), tuple(first) + (0,) * extra))
    return tuple(out)


def _kdim(keep, k):
    k = np.asarray(k)
    return centered(k) if keep else OffsetArray.with_first(np.ones(1, dtype=k.dtype), (0,))


def gradfactors(extended, d, k1, k2):
    """src/kernelfactors.jl:598-602.  `d` is 1-based like the reference's API."""
    N = len(extended)
    return kernelfactors(tuple(_kdim(extended[i], k1 if i == d - 1 else k2) for i in range(N)))


def _pair(f1, f2):
    f1, f2 = centered(np.asarray(f1, dtype=np.float64)), centered(np.asarray(f2, dtype=np.float64))
    return kernelfactors((f2, f1)), kernelfactors((f1, f2))


def _grad(name, k1, k2):
    def fun(extended=None, d=None):
        if extended is None:
            return _pair(k2, k1)
        return gradfactors(tuple(extended), d, np.asarray(k1, dtype=np.float64), np.asarray(k2, dtype=np.float64))
    fun.__name__ = name
    fun.__doc__ = f"KernelFactors.{name}() / {name}(extended, d)  (src/kernelfactors.jl:172-330)"
    return fun


_D = np.array([-1.0, 0.0, 1.0]) / 2
sobel = _grad("sobel", _D, np.array([1.0, 2.0, 1.0]) / 4)
prewitt = _grad("prewitt", _D, np.array([1.0, 1.0, 1.0]) / 3)
scharr = _grad("scharr", _D, np.array([3.0 / 32.0, 5.0 / 16.0, 3.0 / 32.0]) * 2)
bickley = _grad("bickley", _D, np.array([1.0 / 12.0, 1.0 / 3.0, 1.0 / 12.0]) * 2)
ando3 = _grad("ando3", _D, 2 * np.array([0.112737, 0.274526, 0.112737]))


def ando4(extended=None, d=None):
    """src/kernelfactors.jl:332-368 (2-D only)."""
    f1 = np.array([0.0919833, 0.408017, 0.408017, 0.0919833])
    f2 = 1.46205884 * np.array([-0.0919833, -0.408017, 0.408017, 0.0919833])
    pair = _pair(f1, f2)
    if extended is None:
        return pair
    if len(extended) == 2 and all(extended):
        return pair[d - 1]
    raise ArgumentError("dimensions other than 2 are not yet supported")


def ando5(extended=None, d=None):
    """src/kernelfactors.jl:370-404 (2-D only)."""
    f1 = np.array([0.0357338, 0.248861, 0.43081, 0.248861, 0.0357338])
    f2 = 0.784406 * np.array([-0.137424, -0.362576, 0.0, 0.362576, 0.137424])
    pair = _pair(f1, f2)
    if extended is None:
        return pair
    if len(extended) == 2 and all(extended):
        return pair[d - 1]
    raise ArgumentError("dimensions other than 2 are not yet supported")


def box(*sz):
    """src/kernelfactors.jl:172-176."""
    if len(sz) == 1 and isinstance(sz[0], (tuple, list)):
        sz = tuple(sz[0])
    if not all(s % 2 == 1 for s in sz):
        raise ArgumentError(f"kernel dimensions must be odd, got {sz}")
    return kernelfactors(tuple(centered(np.full(s, 1.0 / s)) for s in sz))


def _gaussian1(sigma, l=None):
    """src/kernelfactors.jl:438-443.  eltype follows σ: Python int/float -> Float64, np.float32 -> Float32."""
    if l is None:
        l = 4 * math.ceil(float(sigma)) + 1
    l = int(l)
    if l % 2 != 1:
        raise ArgumentError("length must be odd")
    w = l >> 1
    T = np.float32 if isinstance(sigma, np.float32) else np.float64
    if sigma == 0:
        g = np.array([1.0], dtype=T)
    else:
        x = np.arange(-w, w + 1).astype(T)
        s = T(sigma)
        g = np.exp(-(x * x) / (T(2) * s * s)).astype(T)
    return centered((g / g.sum(dtype=T)).astype(T))


def gaussian(sigma, l=None):
    """gaussian(σ[, l]) -> 1-D factor;  gaussian((σ1, σ2, …)[, (l1, l2, …)]) -> tuple of factors."""
    if isinstance(sigma, (tuple, list, np.ndarray)):
        ls = [None] * len(sigma) if l is None else list(l)
        return kernelfactors(tuple(_gaussian1(s, ll) for s, ll in zip(sigma, ls)))
    return _gaussian1(sigma, l)


def _iirgaussian1(T, sigma, emit_warning=True):
    """IIRGaussian(T, σ)  (src/kernelfactors.jl:533-549; Young, van Vliet & van Ginkel 2002).  The reference forms q in the
    arithmetic of σ mixed with Float64 literals (i.e. Float64), converts it to T and continues in T."""
    import warnings
    if emit_warning and sigma < 1 and sigma != 0:
        warnings.warn("σ is too small for accuracy")
    T = np.dtype(T).type
    sg = float(sigma)
    m0, m1, m2 = T(1.16680), T(1.10783), T(1.40586)
    q = T(1.31564 * (math.sqrt(1 + 0.490811 * sg * sg) - 1))
    two, three, four = T(2), T(3), T(4)
    ascale = (m0 + q) * (m1 * m1 + m2 * m2 + (two * m1) * q + q * q)
    t = (m0 * (m1 * m1 + m2 * m2)) / ascale
    B = t * t
    a1 = (q * ((two * m0) * m1 + m1 * m1 + m2 * m2 + (two * m0 + four * m1) * q + (three * q) * q)) / ascale
    a2 = ((-q * q) * (m0 + two * m1 + three * q)) / ascale
    a3 = ((q * q) * q) / ascale
    return TriggsSdika((a1, a2, a3), B, dtype=T)


def _iirgt(sigma):
    return np.float32 if isinstance(sigma, np.float32) else np.float64


def IIRGaussian(*args, emit_warning=True):
    """IIRGaussian([T], σ) -> TriggsSdika;  IIRGaussian([T], (σ1, σ2, …)) -> tuple of IIR factors, one per dimension
    (src/kernelfactors.jl:517-565).  T defaults to the float type of σ (Float64 for Python numbers)."""
    args = list(args)
    T = None
    if len(args) == 2:
        T, sigma = args
    else:
        (sigma,) = args
    if isinstance(sigma, (tuple, list, np.ndarray)):
        sig = tuple(sigma)
        if T is None:
            T = np.result_type(*[_iirgt(x) for x in sig])
        N = len(sig)
        return tuple(ReshapedOneD(N, d, _iirgaussian1(T, x, emit_warning)) for d, x in enumerate(sig))
    return _iirgaussian1(T if T is not None else _iirgt(sigma), sigma, emit_warning)
