"""Minimal OffsetArray / centered (OffsetArrays.jl semantics used by the reference's kernels).

An `OffsetArray` is a numpy array plus the index of its first element along each axis.  Axis k of
the numpy array is Julia dimension k+1, so `A[i, j]` means the same element in both languages
(modulo the 1-based / offset index origin carried in `first`).
"""
from __future__ import annotations

import numpy as np


class OffsetArray:
    __slots__ = ("parent", "first")

    def __init__(self, parent, *axes):
        """OffsetArray(A, r1, r2, …) with ranges (lo, hi) / range objects, or OffsetArray(A, o1, o2, …)
        with integer OFFSETS (first index = 1 + offset), like OffsetArrays.jl."""
        parent = np.asarray(parent)
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)) and parent.ndim != 1:
            axes = tuple(axes[0])
        if len(axes) != parent.ndim:
            raise ValueError(f"need {parent.ndim} axes/offsets, got {len(axes)}")
        first = []
        for d, ax in enumerate(axes):
            if isinstance(ax, range):
                if len(ax) != parent.shape[d]:
                    raise ValueError("axis length mismatch")
                first.append(ax.start)
            elif isinstance(ax, (tuple, list)):
                lo, hi = ax
                if hi - lo + 1 != parent.shape[d]:
                    raise ValueError("axis length mismatch")
                first.append(int(lo))
            else:
                first.append(1 + int(ax))
        self.parent = parent
        self.first = tuple(first)

    @classmethod
    def with_first(cls, parent, first):
        self = cls.__new__(cls)
        self.parent = np.asarray(parent)
        self.first = tuple(int(f) for f in first)
        return self

    @property
    def ndim(self):
        return self.parent.ndim

    @property
    def shape(self):
        return self.parent.shape

    @property
    def dtype(self):
        return self.parent.dtype

    @property
    def axes(self):
        return tuple(range(f, f + n) for f, n in zip(self.first, self.parent.shape))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return self.parent[tuple(i - f for i, f in zip(idx, self.first))]

    def __setitem__(self, idx, v):
        if not isinstance(idx, tuple):
            idx = (idx,)
        self.parent[tuple(i - f for i, f in zip(idx, self.first))] = v

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.parent, dtype=dtype)

    def __repr__(self):
        return f"OffsetArray(axes={[(a.start, a.stop - 1) for a in self.axes]}, {self.parent!r})"


def centered(a):
    """OffsetArrays.centered: index 0 at the centre element, (first+last)÷2 rounded down."""
    p = a.parent if isinstance(a, OffsetArray) else np.asarray(a)
    return OffsetArray.with_first(p, tuple(-((n - 1) // 2) for n in p.shape))


def parent_and_first(a):
    """(numpy array, first indices) of an OffsetArray or a plain (1-based) array."""
    if isinstance(a, OffsetArray):
        return a.parent, a.first
    a = np.asarray(a)
    return a, (1,) * a.ndim
