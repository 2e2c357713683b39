"""Host-side mirror of the reference API: kernel constructors, typing rules, border specs."""
import math

import numpy as np
import pytest


def test_gaussian_factors_match_survey_appendix_b(ifb):
    g = ifb.KernelFactors.gaussian((3, 3))
    assert len(g) == 2 and g[0].axis == 0 and g[1].axis == 1
    k = g[0].data
    assert k.first == (-6,) and k.shape == (13,) and k.dtype == np.float64
    assert abs(k.parent.sum() - 1) < 1e-15
    np.testing.assert_allclose(k.parent[:7], [0.018544, 0.034167, 0.056332, 0.083109, 0.109719, 0.129618, 0.137023], atol=5e-7)
    g3 = ifb.KernelFactors.gaussian((4, 4, 4))
    assert g3[2].data.first == (-8,) and g3[2].data.shape == (17,)
    assert abs(g3[0].data[0] - 0.103153) < 1e-6 and abs(g3[0].data[-8] - 0.013960) < 1e-6
    assert ifb.KernelFactors.gaussian(np.float32(2)).dtype == np.float32
    with pytest.raises(ifb.ArgumentError):
        ifb.KernelFactors.gaussian(1.0, 4)


def test_log_kernel_facts(ifb):
    k = ifb.Kernel.LoG(3)
    assert k.shape == (27, 27) and k.first == (-13, -13)
    assert abs(k[0, 0] - (-0.0039297517)) < 1e-9
    assert abs(np.abs(k.parent).sum() - 0.16294) < 1e-5
    s = np.linalg.svd(k.parent, compute_uv=False)
    assert s[1] > 1e-3 and s[2] < 1e-12        # exactly rank 2 -> dense path in factorkernel
    fk = ifb.factorkernel(k)
    assert len(fk) == 2 and fk[0].shape == (1, 1) and fk[1].shape == (27, 27)
    # LoG scale golden (reference test/extrema.jl:16-22)
    assert -1.0 * ifb.Kernel.LoG(1.0)[0, 0] == pytest.approx(0.3183098861837907, abs=1e-15)


def test_factorkernel_separable_matrix(ifb):
    g = ifb.Kernel.gaussian(2)
    f = ifb.factorkernel(g)
    assert len(f) == 2 and f[0].shape == (9, 1) and f[1].shape == (1, 9)
    np.testing.assert_allclose(f[0].parent @ f[1].parent, g.parent, atol=1e-15)
    assert f[0].first == (-4, 0) and f[1].first == (0, -4)


def test_sobel_and_gradfactors(ifb):
    k = ifb.KernelFactors.sobel((True, True), 1)
    np.testing.assert_array_equal(k[0].data.parent, [-0.5, 0, 0.5])
    np.testing.assert_array_equal(k[1].data.parent, [0.25, 0.5, 0.25])
    k = ifb.KernelFactors.sobel((True, False, True), 3)
    assert k[1].data.shape == (1,) and k[1].data.first == (0,)
    np.testing.assert_array_equal(k[2].data.parent, [-0.5, 0, 0.5])
    a, b = ifb.KernelFactors.ando4()
    assert a[0].data.first == (-1,) and a[0].data.shape == (4,)     # centered() of an even length
    with pytest.raises(ifb.ArgumentError):
        ifb.KernelFactors.ando4((True, True, True), 1)
    d = ifb.Kernel.sobel()
    np.testing.assert_allclose(d[0].parent, np.outer([-0.5, 0, 0.5], [0.25, 0.5, 0.25]))


def test_reflect(ifb):
    k = ifb.OffsetArray(np.array([[1, 2, 3], [4, 5, 6]]), range(-1, 1), range(0, 3))
    r = ifb.reflect(k)
    assert r.first == (0, -2)
    for i in (-1, 0):
        for j in (0, 1, 2):
            assert r[-i, -j] == k[i, j]


def test_filter_type_rules(ifb):
    ft = ifb.filter_type
    f64k, f32k, ik = np.zeros(3), np.zeros(3, np.float32), np.zeros(3, np.int64)
    assert ft(np.zeros(3, np.float32), f64k) == np.float64
    assert ft(np.zeros(3, np.float32), f32k) == np.float32
    assert ft(np.zeros(3, np.float32), ik) == np.float32
    assert ft(ifb.n0f8(np.zeros(3, np.uint8)), f64k) == np.float64
    assert ft(ifb.n0f8(np.zeros(3, np.uint8)), f32k) == np.float32
    assert ft(np.zeros(3, np.uint8), ik) == np.int64
    assert ft(np.zeros(3, np.uint8), np.zeros(3, np.uint8)) == np.uint8
    assert ft(np.zeros(3, np.int64), f64k) == np.float64
    L = ifb.Kernel.Laplacian()
    assert ft(np.zeros((3, 3), np.uint8), L) == np.int16
    assert ft(np.zeros((3, 3), np.float32), L) == np.float32
    assert ft(np.zeros((3, 3), np.bool_), L) == np.int8
    assert ft(np.zeros(3, np.float32), (f32k, f64k)) == np.float64


def test_border_specs(ifb):
    assert ifb.borderinstance("replicate").style == "replicate"
    with pytest.raises(ifb.ArgumentError):
        ifb.borderinstance("inner")
    with pytest.raises(ifb.ArgumentError):
        ifb.borderinstance("nonsense")
    p = ifb.Pad("circular", (1, 2), (3, 4))
    assert p.lo == (1, 2) and p.hi == (3, 4)
    assert ifb.Pad((1, 1), (2, 2)).style == "replicate"
    assert ifb.Pad("symmetric", (), (1, 1)).lo == (0, 0)
    b = ifb.Fill(7, (1, 1)).to_abi(2)
    assert b.style == ifb._abi.FILL and b.fill == 7.0 and b.npad == 2
    with pytest.raises(ifb.ArgumentError):
        ifb.Pad("reflect", (1, 1, 1), (1, 1, 1)).to_abi(2)


def test_cpu_resources_are_rejected_not_emulated(ifb):
    img = np.zeros((4, 4))
    k = ifb.KernelFactors.gaussian((1, 1))
    for r in (ifb.CPU1(ifb.Algorithm.FIR()), ifb.CPUThreads(ifb.Algorithm.FIRTiled())):
        with pytest.raises(ifb.NotSupportedError):
            ifb.imfilter(r, img, k)
    with pytest.raises(ifb.NotSupportedError):
        ifb.imfilter(ifb.CPU1(ifb.Algorithm.FFT()), img, k)       # FFT() runs on the device (b2f_imfilter_fft), never on the CPU
    with pytest.raises(ifb.NotSupportedError):
        ifb.mapwindow(np.std, img, (3, 3))        # arbitrary window functions cannot cross the C ABI (median / mean / sum can: they are kernels)


def test_n0f8_division_free_conversion_is_correctly_rounded():
    """The device converts N0f8 by q=i*r; rem=fma(-q,255,i); q+=rem*r.  Emulate it here (math.fma needs
    py3.13, so use exact rationals) and compare with IEEE i/255 for all 256 codes, f64 and f32."""
    from fractions import Fraction

    def rn(x, dtype):   # round an exact rational to the nearest float of `dtype`
        if dtype is np.float64:
            return float(x)   # Fraction -> float is correctly rounded
        f = np.float32(float(x))
        # double rounding guard: compare neighbours exactly
        cands = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
        return min(cands, key=lambda c: abs(Fraction(float(c)) - x))

    for dtype in (np.float64, np.float32):
        r = dtype(1) / dtype(255)
        for i in range(256):
            x = dtype(i)
            q = dtype(x * r)
            rem = rn(Fraction(float(x)) - Fraction(float(q)) * 255, dtype)
            v = rn(Fraction(float(rem)) * Fraction(float(r)) + Fraction(float(q)), dtype)
            assert dtype(v) == dtype(x / dtype(255)), (dtype, i)


def test_resolve_window(ifb):
    from importlib import import_module
    mw = import_module("imagefiltering_jl_b200.mapwindow")
    assert mw.resolve_window((3, 5), 2) == ([-1, -2], [1, 2])
    assert mw.resolve_window(3, 1) == ([-1], [1])
    assert mw.resolve_window((range(0, 3), range(-2, 1)), 2) == ([0, -2], [2, 0])
    with pytest.raises(ifb.ArgumentError):
        mw.resolve_window((2, 3), 2)
    assert mw.resolve_window((2, 4), 2, allow_even=True) == ([-1, -2], [0, 1])
    with pytest.raises(ifb.ArgumentError):
        mw.resolve_window((), 0)


def test_color_kernel_lifting(ifb):
    """ColorArray: the kernel factors move one axis up, dense blocks get a leading axis of length 1 (color.py)."""
    from importlib import import_module
    col = import_module("imagefiltering_jl_b200.color")
    k1, k2 = ifb.KernelFactors.gaussian((1, 2))
    l1, l2 = col.lift_kernel((k1, k2))
    assert (l1.N, l1.Npre, l2.N, l2.Npre) == (3, 1, 3, 2) and l1.data is k1.data
    dense = ifb.OffsetArray.with_first(np.arange(6.0).reshape(2, 3), (-1, 0))
    (ld,) = col.lift_kernel((dense,))
    assert ld.parent.shape == (1, 2, 3) and tuple(ld.first) == (0, -1, 0) and np.array_equal(ld.parent[0], dense.parent)
    (ll,) = col.lift_kernel((ifb.Kernel.Laplacian((True, False)),))
    assert tuple(ll.flags) == (False, True, False)
    c = ifb.ColorArray(np.zeros((3, 5, 7), dtype=np.uint8))
    assert (c.nchannels, c.shape, c.ndim) == (3, (5, 7), 2) and c.data.flags.f_contiguous
    with pytest.raises(TypeError):
        ifb.ColorArray(np.zeros(3))


def test_na_border_modes(ifb):
    """NA(na): the three predicates the reference exercises are modes; anything else is refused, not emulated."""
    assert (ifb.NA().mode, ifb.NA("!isfinite").mode, ifb.NA("never").mode) == (0, 1, 2)
    with pytest.raises(ifb.NotSupportedError):
        ifb.NA("x -> x > 3")
    with pytest.raises(ifb.ArgumentError):
        ifb.NA().to_abi(2)            # NA is resolved by imfilter itself, never by the pad


def test_local_extrema_argument_checks(ifb, oracle):
    A = np.zeros((4, 5))
    with pytest.raises(ifb.ArgumentError):
        ifb.findlocalmaxima(A, window=(3,), _library=oracle)
    with pytest.raises(ifb.ArgumentError):
        ifb.findlocalmaxima(A, edges=(True,), _library=oracle)
    with pytest.raises(ifb.ArgumentError):
        ifb.blob_LoG(A, [1.0], edges=(True, False), _library=oracle)
    assert ifb.findlocalmaxima(np.zeros((0, 3)), _library=oracle) == []
    rows = ifb.findlocalmaxima(np.eye(5), as_array=True, _library=oracle)
    assert rows.dtype == np.int64 and rows.shape[1] == 2
    b = ifb.BlobLoG((5, 5), (1.0, 1.0), 0.25)
    assert "CartesianIndex(5, 5)" in repr(b) and b == ifb.BlobLoG((5, 5), (1.0, 1.0), 0.25)


def test_median_window_resolution(ifb, oracle):
    """mapwindow(median!, ...): odd Dims or ranges; even Dims are an ArgumentError like the reference's resolve_window."""
    x = np.arange(10.0)
    with pytest.raises(ifb.ArgumentError):
        ifb.mapwindow(ifb.median, x, (4,), _library=oracle)
    even = ifb.mapwindow(ifb.median, x, range(0, 2), ifb.Fill(0), _library=oracle)      # window i:i+1 -> mean of two neighbours
    assert np.array_equal(even[:-1], x[:-1] / 2 + x[1:] / 2) and even[-1] == 4.5
    with pytest.raises(ifb.NotSupportedError):
        ifb.mapwindow(ifb.median, x, range(1, 3), "replicate", _library=oracle)       # window without its centre: Fill / Inner only
    f32 = ifb.mapwindow(ifb.median, x.astype(np.float32), (3,), _library=oracle)
    assert f32.dtype == np.float32 and ifb.mapwindow(ifb.median, x.astype(np.int32), (3,), _library=oracle).dtype == np.float64
