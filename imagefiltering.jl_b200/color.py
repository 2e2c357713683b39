"""Colour element types (`RGB{N0f8}`, `RGB{Float32}`, `Gray`, any fixed-size colorant) for `imfilter`.

The reference filters colour images through element-type arithmetic (`RGB * Float64`, src/imfilter.jl:1131-1154;
goldens test/2d.jl:49-86,147-226): every channel is filtered independently with the same kernel.  In memory an
`Array{RGB{N0f8},N}` IS a `(3, dims...)` array of bytes, so here a colour image is that array plus a flag: the kernel's
factors are lifted by one leading axis of extent 1 and the ordinary N+1-dimensional path runs — no colour-specific
kernel exists and none is needed.
"""
from __future__ import annotations

import numpy as np

from .kernel import Laplacian
from .kernelfactors import ReshapedOneD
from .offsetarrays import OffsetArray


class ColorArray:
    """`channels`-leading view of a colour image: `data.shape == (C, dims...)`, Fortran order = Julia's memory layout of
    `Array{RGB{T}}`.  uint8 data are N0f8 channels (raw byte i = i/255), float data are float channels."""
    __slots__ = ("data",)

    def __init__(self, data):
        data = np.asarray(data)
        if data.ndim < 2:
            raise TypeError("ColorArray needs a channel axis and at least one spatial axis")
        self.data = np.asfortranarray(data)

    @property
    def nchannels(self):
        return self.data.shape[0]

    @property
    def shape(self):
        return self.data.shape[1:]

    @property
    def ndim(self):
        return self.data.ndim - 1

    def channel(self, c):
        return self.data[c]

    def __repr__(self):
        return f"ColorArray({self.nchannels} channels, {self.shape}, {self.data.dtype})"


def lift_kernel(kernel):
    """Processed kernel tuple for N dims -> the same factors acting on axes 2..N+1 of the channel-leading array."""
    out = []
    for k in kernel:
        if isinstance(k, ReshapedOneD):
            out.append(ReshapedOneD(k.N + 1, k.Npre + 1, k.data))
        elif isinstance(k, Laplacian):
            out.append(Laplacian((False,) + tuple(k.flags)))
        elif isinstance(k, OffsetArray):
            out.append(OffsetArray.with_first(k.parent.reshape((1,) + k.parent.shape, order="F"), (0,) + tuple(k.first)))
        else:
            a = np.asarray(k)
            out.append(OffsetArray.with_first(a.reshape((1,) + a.shape, order="F"), (0,) + (1,) * a.ndim))
    return tuple(out)
