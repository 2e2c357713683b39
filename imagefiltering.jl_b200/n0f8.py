"""N0f8 (FixedPointNumbers.Normed{UInt8,8}) images: raw byte i stands for the value i/255."""
from __future__ import annotations

import numpy as np


class N0f8Array:
    """A uint8 array whose elements are to be read as N0f8 values (`reinterpret(N0f8, raw)`)."""
    __slots__ = ("raw",)

    def __init__(self, raw):
        raw = np.asarray(raw)
        if raw.dtype != np.uint8:
            raise TypeError("N0f8Array wraps a uint8 array")
        self.raw = raw

    @property
    def shape(self):
        return self.raw.shape

    @property
    def ndim(self):
        return self.raw.ndim

    def __array__(self, dtype=None, copy=None):
        return (self.raw.astype(np.float64) / 255.0).astype(dtype or np.float64)


def n0f8(raw):
    return N0f8Array(raw)
