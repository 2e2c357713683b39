"""imagefiltering.jl_b200 — B200-native FIR `imfilter` / running-extrema `mapwindow`, behind the
API of JuliaImages/ImageFiltering.jl.

The product is `libb2f.so` (hand-written CUDA for sm_100a, C ABI in include/b2f.h).  This Python
package is the host-side mirror of the reference's Julia API for that path — the Julia toolchain
is not available in the build image — and carries no arithmetic of its own.

Because the directory name contains a dot, import it through the `imagefiltering_jl_b200` shim
module at the repository root.
"""
from . import _abi, kernel as Kernel, kernelfactors as KernelFactors
from ._abi import (ArgumentError, CudaError, DimensionMismatch, InexactError, NotSupportedError)
from .border import Fill, Inner, NA, NoPad, Pad, borderinstance
from .color import ColorArray
from .device import DeviceArray
from .imfilter import accum_mode, factorkernel, filter_type, imfilter, imfilter_, imgradients, padarray
from .kernel import reflect
from .localextrema import BlobLoG, blob_LoG, findlocalmaxima, findlocalminima
from .kernelfactors import ReshapedOneD, kernelfactors
from .mapwindow import extrema, mapwindow, mapwindow_, maximum, mean, median, median_, minimum, sum_
from .n0f8 import N0f8Array, n0f8
from .offsetarrays import OffsetArray, centered
from .resources import Algorithm, CPU1, CPUThreads, CUDALibs

__all__ = [
    "Kernel", "KernelFactors", "Pad", "Fill", "Inner", "NoPad", "NA", "borderinstance", "imfilter",
    "imfilter_", "imgradients", "padarray", "mapwindow", "mapwindow_", "extrema", "minimum", "maximum", "median", "median_", "mean", "sum_", "centered",
    "OffsetArray", "reflect", "kernelfactors", "ReshapedOneD", "Algorithm", "CUDALibs", "CPU1",
    "CPUThreads", "DeviceArray", "n0f8", "N0f8Array", "filter_type", "factorkernel",
    "ColorArray", "findlocalmaxima", "findlocalminima", "blob_LoG", "BlobLoG", "DimensionMismatch", "ArgumentError", "InexactError", "NotSupportedError", "CudaError",
]
