"""IIR (Triggs-Sdika) filtering, SURVEY §8f rank 4: the reference's own tests for this path (test/triggs.jl, test/basic.jl:22-28)
restated against the CPU oracle (not gpu), and the CUDA kernels against the oracle bit for bit (-m gpu)."""
import warnings

import numpy as np
import pytest


def _kf(ifb):
    return ifb.KernelFactors


def test_triggs_matrix_identity_and_eltypes(ifb):
    """test/triggs.jl:18-22 (Triggs & Sdika Eq. 8: M = I1 + B M A) and test/basic.jl:22-28"""
    for sigma in (3.1, 5, 10.0, 20, 50.0, 100.0):
        k = _kf(ifb).IIRGaussian(sigma)
        A = np.zeros((3, 3)); A[0] = k.a; A[1, 0] = A[2, 1] = 1
        B = np.zeros((3, 3)); B[0] = k.b; B[1, 0] = B[2, 1] = 1
        I1 = np.zeros((3, 3)); I1[0, 0] = 1
        M = np.array(k.M, dtype=np.float64)
        assert np.allclose(M, I1 + B @ M @ A)
    assert _kf(ifb).IIRGaussian(3).dtype == np.float64
    assert _kf(ifb).IIRGaussian(np.float32, 3).dtype == np.float32
    kern = _kf(ifb).IIRGaussian([1, np.float32(2.0)], emit_warning=False)
    assert all(k.data.dtype == np.float64 for k in kern)            # iirgt: Int -> Float64, promoted with Float32
    kern = _kf(ifb).IIRGaussian(np.float32, [np.float32(1), np.float32(2.0)])
    assert all(k.data.dtype == np.float32 for k in kern)
    with pytest.warns(UserWarning):
        _kf(ifb).IIRGaussian(0.5)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        _kf(ifb).IIRGaussian(0.5, emit_warning=False)


def test_triggs_1d_impulses_match_gaussians(ifb, oracle):
    """test/triggs.jl:9-33: unit impulses anywhere on a line of 1000, Fill(0), in place, within 10 % of the gaussian"""
    l = 1000
    for sigma in (3.1, 5, 10.0, 20, 50.0, 100.0):
        kernel = _kf(ifb).IIRGaussian(sigma)
        for c in (1, 2, 3, 5, 10, 20, l >> 1, l - 19, l - 9, l - 4, l - 2, l - 1, l):
            a = np.zeros(l)
            a[c - 1] = 1
            af = np.exp(-((np.arange(1, l + 1) - c) ** 2) / (2 * sigma ** 2)) / (sigma * np.sqrt(2 * np.pi))
            ifb.imfilter_(a, a, (kernel,), ifb.Fill(0), _library=oracle)
            assert np.linalg.norm(a - af) < 0.1 * np.linalg.norm(af), (sigma, c)
    with pytest.raises(ifb.DimensionMismatch):
        ifb.imfilter(np.array([1.0, 2.0]), (_kf(ifb).IIRGaussian(2.0),), _library=oracle)


def test_triggs_images(ifb, oracle):
    """test/triggs.jl:45-90"""
    imgf = np.zeros((5, 7)); imgf[2, 3] = 1
    imgg = np.zeros((5, 7), dtype=np.float32); imgg[2, 3] = 1
    sigma = 5
    x, y = np.arange(-2, 3)[:, None], np.arange(-3, 4)[None, :]
    kernel = _kf(ifb).IIRGaussian((sigma, sigma))
    for img in (imgf, imgg):
        cmp_ = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2)) / (sigma ** 2 * 2 * np.pi)
        img0 = img.copy()
        filt = ifb.imfilter(img, kernel, ifb.Fill(0), _library=oracle)
        assert filt.dtype == np.float64                              # Float64 coefficients (σ is an Int)
        assert np.sum((cmp_ - filt) ** 2) < 0.2 ** 2 * np.sum(cmp_ ** 2)
        assert np.array_equal(img, img0) and not np.array_equal(filt, img)
    ret = ifb.imfilter(imgf, _kf(ifb).IIRGaussian((sigma, 0)), "replicate", _library=oracle)
    assert not np.array_equal(ret, imgf)
    ret = ifb.imfilter(imgf, _kf(ifb).IIRGaussian((0, sigma)), "replicate", _library=oracle)
    assert not np.array_equal(ret, imgf)
    out = np.empty_like(ret, order="F")
    ifb.imfilter_(ifb.CUDALibs(ifb.Algorithm.IIR()), out, imgf, _kf(ifb).IIRGaussian(sigma), 2, "replicate", _library=oracle)
    assert np.array_equal(out, ret)
    kerng = _kf(ifb).IIRGaussian(sigma)
    ifb.imfilter(imgf, (kerng, kerng), ifb.NA(), _library=oracle)
    imgfnan, imgfnum = imgf.copy(), imgf.copy()
    imgfnan[0, 0], imgfnum[0, 0] = np.nan, 0
    imgfden = np.ones((5, 7)); imgfden[0, 0] = 0
    retnum = ifb.imfilter(imgfnum, kernel, ifb.Fill(0.0), _library=oracle)
    retden = ifb.imfilter(imgfden, kernel, ifb.Fill(0.0), _library=oracle)
    ret = ifb.imfilter(imgfnan, kernel, ifb.NA(), _library=oracle)
    ret[0, 0] = retnum[0, 0] = 0
    assert np.allclose(ret, retnum / retden)
    with pytest.raises(ifb.ArgumentError):
        ifb.imfilter(imgf, kernel, "reflect", _library=oracle)     # only "replicate" is supported, src/imfilter.jl:897


def test_triggs_offsetarrays(ifb, oracle):
    """test/triggs.jl:93-104"""
    A = np.arange(1, 100 * 100 + 1).reshape((100, 100), order="F")
    kern = tuple(_kf(ifb).IIRGaussian(5.0) for _ in range(2))
    B = ifb.imfilter(A, kern, ifb.NA(), _library=oracle)
    C = ifb.OffsetArray.with_first(A, (0, 0))
    D = ifb.imfilter(C, kern, ifb.NA(), _library=oracle)
    assert np.array_equal(np.asarray(B), D.parent)


def test_mixed_fir_iir_is_rejected(ifb, oracle):
    k1 = _kf(ifb).IIRGaussian(2)
    k2 = ifb.centered(np.ones(3) / 3)
    with pytest.raises(ifb.NotSupportedError):
        ifb.imfilter(np.arange(1.0, 9.0), (k1, k2), _library=oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "f64", "n0f8", "i32"])
@pytest.mark.parametrize("border", ["replicate", "fill"])
def test_gpu_iir_bit_exact(ifb, oracle, device, dt, border):
    """b2f_iir on the GPU (one thread per line; the panel kernel along the contiguous axis) against the oracle: every operation
    is a separate multiply / add in the reference's order, so Float32 and Float64 results are bit-equal."""
    rng = np.random.default_rng({"f32": 1, "f64": 2, "n0f8": 3, "i32": 4}[dt] + (10 if border == "fill" else 0))
    b = ifb.Fill(0.25 if dt in ("f32", "f64") else 0) if border == "fill" else "replicate"
    for shape, sig in (((500,), (7.5,)), ((4,), (2.0,)), ((131, 77), (3.0, 5.0)), ((64, 200), (10.0, 0)), ((33, 21, 18), (2.0, 3.0, 4.0)),
                       ((70, 9, 5, 4), (3.0, 2.0, 0, 1.5))):
        if dt == "f32":
            img = np.asfortranarray(rng.random(shape, dtype=np.float32))
            kern = ifb.KernelFactors.IIRGaussian(tuple(np.float32(s) for s in sig), emit_warning=False)
            want = np.float32
        elif dt == "f64":
            img = np.asfortranarray(rng.random(shape))
            kern = ifb.KernelFactors.IIRGaussian(sig, emit_warning=False)
            want = np.float64
        elif dt == "n0f8":
            img = ifb.n0f8(np.asfortranarray(rng.integers(0, 256, size=shape, dtype=np.uint8)))
            kern = ifb.KernelFactors.IIRGaussian(sig, emit_warning=False)
            want = np.float64
        else:
            img = np.asfortranarray(rng.integers(-1000, 1000, size=shape).astype(np.int32))
            kern = ifb.KernelFactors.IIRGaussian(tuple(np.float32(s) for s in sig), emit_warning=False)
            want = np.float32
        a = ifb.imfilter(img, kern, b)
        assert device.last_path() in ("iir_rows", "iir_strided")
        o = ifb.imfilter(img, kern, b, _library=oracle)
        assert a.dtype == want and o.dtype == want
        assert np.array_equal(a, o), (shape, dt, border)
    # in place, and along one chosen dimension
    img = np.asfortranarray(rng.random((90, 120)))
    k = ifb.KernelFactors.IIRGaussian(6.0)
    x, y = img.copy(order="F"), img.copy(order="F")
    ifb.imfilter_(x, x, (k, k), b)
    ifb.imfilter_(y, y, (k, k), b, _library=oracle)
    assert np.array_equal(x, y)
    for dim in (1, 2):
        x, y = np.empty_like(img, order="F"), np.empty_like(img, order="F")
        ifb.imfilter_(ifb.CUDALibs(ifb.Algorithm.IIR()), x, img, k, dim, b)
        ifb.imfilter_(ifb.CUDALibs(ifb.Algorithm.IIR()), y, img, k, dim, b, _library=oracle)
        assert np.array_equal(x, y)
    # NA border, with and without NaNs
    nan = img.copy(order="F")
    for im in (img, nan):
        nan[3, 5] = np.nan
        a = ifb.imfilter(im, (k, k), ifb.NA())
        o = ifb.imfilter(im, (k, k), ifb.NA(), _library=oracle)
        assert np.array_equal(a, o, equal_nan=True)


def test_triggs_colour_image(ifb, oracle):
    """test/triggs.jl:45-60, `imgc = fill(RGB{Float64}(0,0,0), 5, 7); imgc[3,4] = RGB(1,0,0)`: the red channel is the filtered
    impulse, the others stay zero, the input is untouched."""
    data = np.zeros((3, 5, 7))
    data[0, 2, 3] = 1.0
    img = ifb.ColorArray(data)
    sigma = 5
    x, y = np.arange(-2, 3)[:, None], np.arange(-3, 4)[None, :]
    cmp_ = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2)) / (sigma ** 2 * 2 * np.pi)
    filt = ifb.imfilter(img, ifb.KernelFactors.IIRGaussian((sigma, sigma)), ifb.Fill(0), _library=oracle)
    assert isinstance(filt, ifb.ColorArray) and filt.data.shape == (3, 5, 7)
    assert np.sum((cmp_ - filt.data[0]) ** 2) < 0.2 ** 2 * np.sum(cmp_ ** 2)
    assert not filt.data[1:].any() and data[0, 2, 3] == 1.0
    gray = ifb.imfilter(np.ascontiguousarray(data[0]), ifb.KernelFactors.IIRGaussian((sigma, sigma)), ifb.Fill(0), _library=oracle)
    assert np.array_equal(filt.data[0], gray)
