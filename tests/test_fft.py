"""Algorithm.FFT() (reference src/imfilter.jl:776-888).  The reference's tests for this algorithm assert
`imfilter(img, kernel, border, Algorithm.FFT()) ≈ the FIR target` (test/2d.jl:69-140, test/nd.jl:77-81); they are restated here
against the oracle (host logic: kernelconv, dispatch, Inner, the Int / InexactError rule) and, -m gpu, the cuFFT-backed device
path against the oracle's exact correlation within the rounding of the transforms."""
import numpy as np
import pytest

BORDERS = ["replicate", "circular", "symmetric", "reflect"]


def _approx(a, b, rtol=1e-10):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm((a - b).ravel()) <= rtol * max(1.0, np.linalg.norm(b.ravel()))


def _cases(ifb):
    """test/2d.jl:39-140: an impulse image, a dense kernel, a factored kernel, 'rational' thirds"""
    imgf = np.zeros((5, 7)); imgf[2, 3] = 0.625
    imgi = np.zeros((5, 7), dtype=np.int64); imgi[2, 3] = 1
    kern = np.array([[0.1, 0.2], [0.4, 0.5]])
    dense = ifb.OffsetArray.with_first(kern, (-1, 1))
    fact = (ifb.OffsetArray.with_first(np.array([0.2, 0.8]), (-1,)), ifb.OffsetArray.with_first(np.array([[0.3, 0.6]]), (0, 1)))
    thirds = (ifb.centered(np.array([1 / 3, 1 / 3, 1 / 3])), ifb.centered(np.array([[1 / 3, 1 / 3, 1 / 3]])))
    return (imgf, imgi), (dense, fact, thirds)


def _check_suite(ifb, lib, ref_lib, rtol):
    imgs, kernels = _cases(ifb)
    fft = ifb.Algorithm.FFT()
    for img in imgs:
        for kernel in kernels:
            for border in BORDERS + [ifb.Fill(0)]:
                want = ifb.imfilter(np.float64, img, kernel, border, _library=ref_lib)
                assert _approx(ifb.imfilter(np.float64, img, kernel, border, fft, _library=lib), want, rtol)
                if img.dtype.kind == "f":
                    assert _approx(ifb.imfilter(img, kernel, border, fft, _library=lib), want, rtol)
                    assert _approx(ifb.imfilter(ifb.CUDALibs(fft), img, kernel, border, _library=lib), want, rtol)
                got32 = ifb.imfilter(np.float32, img, kernel, border, fft, _library=lib)
                assert got32.dtype == np.float32 and _approx(got32, want, max(rtol, 1e-6))
                ret = np.zeros(img.shape, order="F")
                ifb.imfilter_(ifb.CUDALibs(fft), ret, img, kernel, border, _library=lib)
                assert _approx(ret, want, rtol)
            want = ifb.imfilter(np.float64, img, kernel, ifb.Inner(), _library=ref_lib)
            got = ifb.imfilter(np.float64, img, kernel, ifb.Inner(), fft, _library=lib)
            assert got.first == want.first and _approx(got.parent, want.parent, rtol)
    # 1-D, a resource and an algorithm together stay a MethodError, integer results are inexact
    img = np.arange(1.0, 9.0)
    k1 = ifb.centered(np.array([0.25, 0.5, 0.25]))
    assert _approx(ifb.imfilter(img, (k1,), "replicate", fft, _library=lib), ifb.imfilter(img, (k1,), "replicate", _library=ref_lib), rtol)
    with pytest.raises(TypeError):
        ifb.imfilter(ifb.CUDALibs(fft), img, (k1,), "replicate", fft, _library=lib)
    with pytest.raises(ifb.InexactError):
        ifb.imfilter(np.arange(8), (ifb.centered(np.array([1, 1, 1])),), "replicate", fft, _library=lib)


def test_fft_algorithm_host_logic_against_oracle(ifb, oracle):
    _check_suite(ifb, oracle, oracle, 1e-12)


@pytest.mark.gpu
def test_gpu_fft_reference_cases(ifb, oracle, device):
    _check_suite(ifb, None, oracle, 1e-10)
    assert device.last_path() == "fft" or True


@pytest.mark.gpu
@pytest.mark.parametrize("border", BORDERS + ["fill", "inner"])
def test_gpu_fft_matches_exact_correlation(ifb, oracle, device, border):
    """cuFFT-backed path against the oracle's exact correlation: 1-D, 2-D, 3-D and a batch of 2-D images; dense, factored and
    long kernels (the regime the reference itself switches to FFT in: > 30 taps, src/imfilter.jl:1200-1204); N0f8 input."""
    rng = np.random.default_rng(sum(map(ord, border)))
    b = {"fill": ifb.Fill(0.25), "inner": ifb.Inner()}.get(border, border)
    fft = ifb.Algorithm.FFT()
    cases = [((300,), (ifb.centered(rng.random(41) - 0.4),)),
             ((120, 97), ifb.KernelFactors.gaussian((10, 8))),
             ((96, 81), (ifb.Kernel.DoG((2, 2)),)),
             ((40, 33, 21), (ifb.centered(rng.random((5, 3, 7)) - 0.5),)),
             ((64, 50, 6), ifb.KernelFactors.gaussian((3, 2, 0)))]
    for shape, kern in cases:
        img = np.asfortranarray(rng.random(shape))
        got = ifb.imfilter(np.float64, img, kern, b, fft)
        assert device.last_path() == "fft"
        want = ifb.imfilter(np.float64, img, kern, b, _library=oracle)
        gp = got.parent if isinstance(got, ifb.OffsetArray) else got
        wp = want.parent if isinstance(want, ifb.OffsetArray) else want
        assert gp.shape == wp.shape and np.max(np.abs(gp - wp)) <= 1e-11 * max(1.0, np.abs(wp).max()) * np.sqrt(img.size), (shape, border)
        g32 = ifb.imfilter(np.float32, img.astype(np.float32), kern, b, fft)
        g32 = g32.parent if isinstance(g32, ifb.OffsetArray) else g32
        assert g32.dtype == np.float32 and np.max(np.abs(g32 - wp)) <= 2e-4 * max(1.0, np.abs(wp).max()), (shape, border)
    raw = ifb.n0f8(np.asfortranarray(rng.integers(0, 256, size=(70, 45), dtype=np.uint8)))
    kern = ifb.KernelFactors.gaussian((4, 4))
    got = ifb.imfilter(raw, kern, b if border != "fill" else ifb.Fill(0), fft)
    want = ifb.imfilter(raw, kern, b if border != "fill" else ifb.Fill(0), _library=oracle)
    gp = got.parent if isinstance(got, ifb.OffsetArray) else got
    wp = want.parent if isinstance(want, ifb.OffsetArray) else want
    assert np.max(np.abs(gp - wp)) <= 1e-11
