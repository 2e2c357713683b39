"""Algorithm tags (reference src/ImageFiltering.jl:52-68) and ComputationalResources-style resources.

`CUDALibs(Algorithm.FIR())` is the resource this package implements; it is the seam the reference
itself advertises (src/imfilter.jl:63-73, src/ImageFiltering.jl:105-112).  CPU resources exist only
so that calls written for the reference fail with a clear message instead of silently running on
the CPU: this package has no CPU execution path.
"""
from __future__ import annotations


class Alg:
    pass


class FIR(Alg):
    def __repr__(self):
        return "Algorithm.FIR()"


class FIRTiled(Alg):
    def __init__(self, tilesize=()):
        self.tilesize = tuple(tilesize)

    def __repr__(self):
        return f"Algorithm.FIRTiled({self.tilesize})"


class FFT(Alg):
    pass


class IIR(Alg):
    pass


class Mixed(Alg):
    pass


class Algorithm:
    Alg = Alg
    FIR = FIR
    FIRTiled = FIRTiled
    FFT = FFT
    IIR = IIR
    Mixed = Mixed


class AbstractResource:
    def __init__(self, settings=None):
        self.settings = settings

    def __repr__(self):
        return f"{type(self).__name__}({self.settings!r})"


class CPU1(AbstractResource):
    pass


class CPUThreads(AbstractResource):
    pass


class CUDALibs(AbstractResource):
    """The B200 resource.  `settings` is an Algorithm.FIR() (default)."""

    def __init__(self, settings=None):
        super().__init__(FIR() if settings is None else settings)
