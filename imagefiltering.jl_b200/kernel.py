"""Kernel — dense kernel constructors (reference src/kernel.jl).  Host-side data producers."""
from __future__ import annotations

import math

import numpy as np

from . import kernelfactors as KF
from ._abi import ArgumentError
from .offsetarrays import OffsetArray, centered


def _bcast_product(factors):
    """broadcast(*, factors...) of ReshapedOneD / OffsetArray factors."""
    dens = [f.dense() if isinstance(f, KF.ReshapedOneD) else f for f in factors]
    p = dens[0].parent
    for d in dens[1:]:
        p = p * d.parent
    first = tuple(next((d.first[ax] for d in dens if d.shape[ax] > 1), dens[0].first[ax])
                  for ax in range(p.ndim))
    return OffsetArray.with_first(p, first)


def _product2d(kf):
    k1, k2 = kf
    return _bcast_product(k1), _bcast_product(k2)


def _gradfamily(name):
    kf = getattr(KF, name)

    def fun(extended=None, d=None):
        if extended is None:
            return _product2d(kf())
        return (_bcast_product(kf(tuple(extended), d)),)
    fun.__name__ = name
    fun.__doc__ = f"Kernel.{name}() / {name}(extended, d)  (src/kernel.jl:44-230)"
    return fun


sobel = _gradfamily("sobel")
prewitt = _gradfamily("prewitt")
scharr = _gradfamily("scharr")
bickley = _gradfamily("bickley")
ando3 = _gradfamily("ando3")


def ando4(extended=None, d=None):
    f = np.array([[-0.022116, -0.025526, 0.025526, 0.022116],
                  [-0.098381, -0.112984, 0.112984, 0.098381],
                  [-0.098381, -0.112984, 0.112984, 0.098381],
                  [-0.022116, -0.025526, 0.025526, 0.022116]])
    pair = (centered(f.T.copy()), centered(f))
    if extended is None:
        return pair
    if not all(extended):
        raise ArgumentError("all dimensions must be extended")
    return (pair[d - 1],)


def ando5(extended=None, d=None):
    f = np.array([[-0.003776, -0.010199, 0.0, 0.010199, 0.003776],
                  [-0.026786, -0.070844, 0.0, 0.070844, 0.026786],
                  [-0.046548, -0.122572, 0.0, 0.122572, 0.046548],
                  [-0.026786, -0.070844, 0.0, 0.070844, 0.026786],
                  [-0.003776, -0.010199, 0.0, 0.010199, 0.003776]])
    pair = (centered(f.T.copy()), centered(f))
    if extended is None:
        return pair
    if not all(extended):
        raise ArgumentError("all dimensions must be extended")
    return (pair[d - 1],)


def box(sz):
    return _bcast_product(KF.box(tuple(sz)))


def gaussian(sigma, l=None):
    """src/kernel.jl:232-262: scalar σ means isotropic 2-D."""
    if not isinstance(sigma, (tuple, list, np.ndarray)):
        sigma = (sigma, sigma)
    if len(sigma) == 1:
        return KF.gaussian(sigma[0], None if l is None else l[0])
    return _bcast_product(KF.gaussian(tuple(sigma), l))


def DoG(sigma_p, sigma_m=None, ls=None):
    """src/kernel.jl:264-300."""
    if not isinstance(sigma_p, (tuple, list)):
        sigma_p = (sigma_p, sigma_p)
    if sigma_m is None:
        sigma_m = tuple(s * math.sqrt(2) for s in sigma_p)
        neg = gaussian(tuple(sigma_m))
        ls = neg.shape
    else:
        neg = gaussian(tuple(sigma_m), ls)
    pos = gaussian(tuple(sigma_p), ls)
    return OffsetArray.with_first(pos.parent - neg.parent, pos.first)


def LoG(sigma):
    """src/kernel.jl:325-339: half-width ceil(8.5σ)>>1 per axis."""
    if not isinstance(sigma, (tuple, list)):
        sigma = (sigma, sigma)
    N = len(sigma)
    ws = [int(math.ceil(8.5 * s)) >> 1 for s in sigma]
    s = np.asarray(sigma, dtype=np.float64)
    Cn = 1.0 / (np.prod(s) * (2 * math.pi) ** (N / 2))
    s2 = s ** 2
    s2i = np.sum(1.0 / s2)
    grids = np.meshgrid(*[np.arange(-w, w + 1, dtype=np.float64) for w in ws], indexing="ij")
    xs = [g ** 2 / s2[d] for d, g in enumerate(grids)]
    sum_xs_s = sum(x / s2[d] for d, x in enumerate(xs))
    sum_xs = sum(xs)
    out = Cn * ((sum_xs_s - s2i) * np.exp(-sum_xs / 2))
    return OffsetArray.with_first(out, tuple(-w for w in ws))


class Laplacian:
    """Kernel.Laplacian((true,true,…)) (src/kernel.jl:341-391): an opaque stencil type."""

    def __init__(self, flags=(True, True), N=None):
        if N is not None:  # Laplacian(dims, N), dims 1-based
            fl = [False] * N
            for d in flags:
                fl[d - 1] = True
            flags = fl
        self.flags = tuple(bool(f) for f in flags)

    @property
    def ndim(self):
        return len(self.flags)

    def asarray(self):
        """convert(AbstractArray, L) (src/kernel.jl:377-385)."""
        shape = [3 if f else 1 for f in self.flags]
        A = np.zeros(shape, dtype=np.int64)
        c = tuple(1 if f else 0 for f in self.flags)
        for d, f in enumerate(self.flags):
            if f:
                for s in (-1, 1):
                    i = list(c)
                    i[d] += s
                    A[tuple(i)] = 1
        A[c] = -2 * sum(self.flags)
        return OffsetArray.with_first(A, tuple(-1 if f else 0 for f in self.flags))


def laplacian2d(alpha=0):
    lc = alpha / (1 + alpha)
    lb = (1 - alpha) / (1 + alpha)
    lm = -4 / (1 + alpha)
    return centered(np.array([[lc, lb, lc], [lb, lm, lb], [lc, lb, lc]], dtype=np.float64))


def gabor(size_x, size_y, sigma, theta, lam, gamma, psi):
    """src/kernel.jl:421-468."""
    if not (sigma > 0 and lam > 0 and gamma > 0):
        raise ArgumentError("The parameters σ, λ and γ must be positive numbers.")
    sx, sy = sigma, sigma / gamma
    c, s = math.cos(theta), math.sin(theta)
    xmax = size_x // 2 if size_x > 0 else int(round(max(abs(3 * sx * c), abs(3 * sy * s), 1)))
    ymax = size_y // 2 if size_y > 0 else int(round(max(abs(3 * sx * s), abs(3 * sy * c), 1)))
    ii, jj = np.meshgrid(np.arange(-xmax, xmax + 1), np.arange(-ymax, ymax + 1), indexing="ij")
    x, y = jj.astype(np.float64), ii.astype(np.float64)
    xr = x * c + y * s
    yr = -x * s + y * c
    env = np.exp(-0.5 * ((xr * xr) / sx ** 2 + (yr * yr) / sy ** 2))
    return env * np.cos(2 * (math.pi / lam) * xr + psi), env * np.sin(2 * (math.pi / lam) * xr + psi)


def moffat(alpha, beta, ls=None):
    """src/kernel.jl:470-511."""
    if ls is None:
        ls = int(math.ceil((alpha * 2 * math.sqrt(2 ** (1 / beta) - 1)) * 4))
    if not isinstance(ls, (tuple, list)):
        ls = (ls, ls)
    ws = [int(math.ceil(n)) >> 1 for n in ls]
    grids = np.meshgrid(*[np.arange(-w, w + 1, dtype=np.float64) for w in ws], indexing="ij")
    r2 = sum(g ** 2 for g in grids)
    a2 = alpha ** 2
    amp = (beta - 1) / (math.pi * a2)
    return OffsetArray.with_first(amp * ((1 + r2 / a2) ** -beta), tuple(-w for w in ws))


def reflect(kernel):
    """src/kernel.jl:520-529: reflect(kernel)[-I] = kernel[I]  (correlation <-> convolution)."""
    if isinstance(kernel, KF.ReshapedOneD):
        return KF.ReshapedOneD(kernel.N, kernel.Npre, reflect(kernel.data))
    if isinstance(kernel, tuple):
        return tuple(reflect(k) for k in kernel)
    p, first = (kernel.parent, kernel.first) if isinstance(kernel, OffsetArray) else (np.asarray(kernel), (1,) * np.asarray(kernel).ndim)
    flipped = p[tuple(slice(None, None, -1) for _ in range(p.ndim))].copy()
    return OffsetArray.with_first(flipped, tuple(-(f + n - 1) for f, n in zip(first, p.shape)))
