"""Differential parity: CUDA library vs the CPU oracle on identical seeded inputs, through the C ABI.

Bar (BASELINE.json north_star): bit-exact for Float64 outputs whose axes are filtered at most once,
for integer kernels on integer data and for min/max; |gpu - oracle| <= 1e-5 * prod_stage(sum|k|) *
max|img| for Float32 outputs.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def seed_of(*key):
    """Reproducible seed from a test's parameters (Python's hash() of strings changes per process)."""
    import zlib
    return zlib.crc32(repr(key).encode())

BORDERS = ["replicate", "circular", "symmetric", "reflect"]


def _tol(stages_taps, img):
    s = 1.0
    for t in stages_taps:
        s *= np.abs(np.asarray(t, dtype=np.float64)).sum()
    return 1e-5 * s * float(np.abs(np.asarray(img, dtype=np.float64)).max())


def _both(ifb, oracle, *args):
    a = ifb.imfilter(*args)
    b = ifb.imfilter(*args, _library=oracle)
    pa = a.parent if isinstance(a, ifb.OffsetArray) else a
    pb = b.parent if isinstance(b, ifb.OffsetArray) else b
    if isinstance(a, ifb.OffsetArray):
        assert a.first == b.first
    assert pa.shape == pb.shape and pa.dtype == pb.dtype
    return pa, pb


@pytest.mark.parametrize("border", BORDERS + ["fill", "inner"])
@pytest.mark.parametrize("shape", [(37, 53), (130, 67), (5, 4), (257, 129, 3)])
@pytest.mark.parametrize("dt", ["f32", "f64", "n0f8", "u8"])
def test_separable_f64_bit_exact(ifb, oracle, device, border, shape, dt):
    rng = np.random.default_rng(seed_of((border, shape, dt)))
    if dt == "f32":
        img = rng.random(shape, dtype=np.float32)
    elif dt == "f64":
        img = rng.random(shape)
    else:
        raw = rng.integers(0, 256, size=shape, dtype=np.uint8)
        img = ifb.n0f8(raw) if dt == "n0f8" else raw
    nd = len(shape)
    b = {"fill": ifb.Fill(0.25 if dt != "u8" else 3), "inner": ifb.Inner()}.get(border, border)
    for sig in ((1, 2), (3, 1)):
        kf = ifb.KernelFactors.gaussian(sig + (0,) * (nd - 2)) if nd > 2 else ifb.KernelFactors.gaussian(sig)
        for kern in (kf, tuple(reversed(kf))):   # x-first and y-first cascades
            pa, pb = _both(ifb, oracle, img, kern, b)
            assert pa.dtype == np.float64
            assert np.array_equal(pa, pb), (border, shape, dt, sig)
            assert device.last_path() in (("stream2d", "fused2d") if pa.size else ("empty",))


@pytest.mark.parametrize("border", BORDERS + ["fill"])
@pytest.mark.parametrize("shape", [(64, 64), (131, 77), (300, 9), (70, 66, 4)])
def test_separable_f32_tolerance(ifb, oracle, device, border, shape):
    rng = np.random.default_rng(seed_of((border, shape)))
    img = rng.random(shape, dtype=np.float32)
    b = ifb.Fill(0.5) if border == "fill" else border
    nd = len(shape)
    for sig in ((3, 3), (1, 6)):
        kf = ifb.KernelFactors.gaussian(sig + (0,) * (nd - 2)) if nd > 2 else ifb.KernelFactors.gaussian(sig)
        pa, pb = _both(ifb, oracle, np.float32, img, kf, b)
        assert pa.dtype == np.float32
        tol = _tol([k.data.parent for k in kf], img)
        assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= tol
        assert device.last_path() in ("stream2d", "fused2d", "sepnd")
        # float32 taps too (KernelFactors.gaussian(σ::Float32))
        kf32 = ifb.KernelFactors.gaussian(tuple(np.float32(s) for s in sig) + (np.float32(0),) * (nd - 2))
        pa, pb = _both(ifb, oracle, img, kf32, b)
        assert pa.dtype == np.float32
        assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= tol


@pytest.mark.parametrize("border", BORDERS + ["fill", "inner"])
@pytest.mark.parametrize("fun", ["sobel", "prewitt", "scharr", "ando5"])
def test_imgradients_n0f8_bit_exact(ifb, oracle, device, border, fun):
    rng = np.random.default_rng(17)
    raw = rng.integers(0, 256, size=(203, 91), dtype=np.uint8)
    img = ifb.n0f8(raw)
    b = {"fill": ifb.Fill(0.0), "inner": ifb.Inner()}.get(border, border)
    kfun = getattr(ifb.KernelFactors, fun)
    ga = ifb.imgradients(img, kfun, b)
    assert device.last_path() in ("stream2d_grad", "fused2d_grad")
    gb = ifb.imgradients(img, kfun, b, _library=oracle)
    for a, o in zip(ga, gb):
        pa = a.parent if isinstance(a, ifb.OffsetArray) else a
        po = o.parent if isinstance(o, ifb.OffsetArray) else o
        assert pa.dtype == np.float64 and np.array_equal(pa, po)
    g32 = ifb.imgradients(img, kfun, b, T=np.float32)
    for a, o in zip(g32, gb):
        pa = a.parent if isinstance(a, ifb.OffsetArray) else a
        po = o.parent if isinstance(o, ifb.OffsetArray) else o
        assert pa.dtype == np.float32 and np.max(np.abs(pa - po)) <= 1e-5


def test_generic_path_cases(ifb, oracle, device):
    """Everything the fused kernels do not cover goes through the per-stage device path."""
    rng = np.random.default_rng(23)
    cases = []
    a1 = rng.random(50)
    k1 = ifb.centered(rng.random(5))
    cases.append((a1, (k1, k1, k1)))                                    # 1-D, same axis three times
    a2 = np.asfortranarray(rng.random((33, 21)))
    kx = ifb.OffsetArray(rng.random((3, 1)), range(-1, 2), range(0, 1))
    ky = ifb.OffsetArray(rng.random((1, 4)), range(0, 1), range(-2, 2))
    cases.append((a2, (kx, ky, kx, ky)))                                # repeated axes (test/cascade.jl)
    cases.append((a2, (ky,)))                                           # single 1-D stage
    cases.append((a2, ifb.OffsetArray(rng.random((3, 4)), range(-1, 2), range(0, 4))))   # dense, asymmetric
    a3 = np.asfortranarray(rng.random((12, 9, 7)))
    cases.append((a3, ifb.KernelFactors.gaussian((1, 1, 1))))           # 3-D separable
    cases.append((a3, ifb.centered(rng.random((3, 3, 3)))))             # 3-D dense
    a4 = np.asfortranarray(rng.random((6, 5, 4, 3)))
    cases.append((a4, ifb.centered(rng.random((3, 1, 3, 1)))))          # 4-D
    for img, kern in cases:
        for border in BORDERS + [ifb.Fill(0.3), ifb.Inner()]:
            pa, pb = _both(ifb, oracle, img, kern, border)
            assert np.array_equal(pa, pb), (img.shape, border)
            assert device.last_path() in ("generic", "dense2d", "dense3d", "sepnd", "fused2d", "stream2d")


def test_integer_exact_and_inexact(ifb, oracle, device):
    rng = np.random.default_rng(29)
    img = rng.integers(0, 256, size=(40, 31), dtype=np.uint8)
    kern = ifb.centered(rng.integers(-3, 4, size=(3, 5)).astype(np.int64))
    pa, pb = _both(ifb, oracle, img, kern, "reflect")
    assert pa.dtype == np.int64 and np.array_equal(pa, pb)
    pa, pb = _both(ifb, oracle, np.int32, img, kern, "circular")
    assert pa.dtype == np.int32 and np.array_equal(pa, pb)
    with pytest.raises(ifb.InexactError):
        ifb.imfilter(np.uint8, img, kern)
    i16 = rng.integers(-1000, 1000, size=(25, 25)).astype(np.int16)
    ksep = (ifb.centered(np.array([1, 2, 1], dtype=np.int64)), ifb.centered(np.array([[1, 0, -1]], dtype=np.int64)))
    pa, pb = _both(ifb, oracle, i16, ksep, "symmetric")
    assert pa.dtype == np.int64 and np.array_equal(pa, pb)


def test_pad_larger_than_image_and_explicit_pads(ifb, oracle, device):
    rng = np.random.default_rng(31)
    img = np.asfortranarray(rng.random((3, 2)))
    kf = ifb.KernelFactors.gaussian((2, 2))          # 9 taps on a 3x2 image: multi-fold remap
    for border in ["replicate", "circular", "symmetric", "reflect", ifb.Fill(1.5)]:
        pa, pb = _both(ifb, oracle, img, kf, border)
        assert np.array_equal(pa, pb), border
    big = np.asfortranarray(rng.random((20, 20)))
    pa, pb = _both(ifb, oracle, big, kf, ifb.Pad("reflect", (6, 6), (5, 5)))   # more than needed: fine
    assert np.array_equal(pa, pb)
    with pytest.raises(ifb.DimensionMismatch):
        ifb.imfilter(big, kf, ifb.Pad("reflect", (1, 1), (1, 1)))              # less than needed
    with pytest.raises(ifb.ArgumentError):
        ifb.imfilter(np.asfortranarray(rng.random((1, 5))), kf, "reflect")     # DivideError in the reference


def test_imfilter_inplace_roi_and_nopad(ifb, oracle, device):
    rng = np.random.default_rng(37)
    img = np.asfortranarray(rng.random((30, 40)))
    kf = ifb.KernelFactors.gaussian((1, 1))
    inds = (range(4, 20), range(6, 30))
    outs = []
    for lib in (None, oracle):
        out = np.full((30, 40), -7.0, order="F")
        ifb.imfilter_(ifb.CUDALibs(), out, img, kf, ifb.NoPad(), inds, _library=lib)
        outs.append(out)
    assert np.array_equal(outs[0], outs[1])
    assert np.all(outs[0][:3, :] == -7.0) and np.all(outs[0][19:, :] == -7.0)
    with pytest.raises(ifb.DimensionMismatch):
        ifb.imfilter_(ifb.CUDALibs(), np.zeros((30, 40), order="F"), img, kf, ifb.NoPad())


def test_device_resident_arrays(ifb, oracle, device):
    """Zero-copy form: torch CUDA tensors wrapped as DeviceArray, batch of images in one launch."""
    import torch
    rng = np.random.default_rng(41)
    B, H, W = 5, 70, 200
    host = rng.integers(0, 256, size=(B, H, W), dtype=np.uint8)          # C-order (B,H,W) == Julia dims (W,H,B)
    t = torch.from_numpy(host).cuda()
    gx = torch.empty((B, H, W), dtype=torch.float64, device="cuda")
    gy = torch.empty_like(gx)
    img_d = ifb.DeviceArray.from_torch(t, n0f8=True)
    k1 = ifb.KernelFactors.sobel((True, True, False), 1)
    k2 = ifb.KernelFactors.sobel((True, True, False), 2)
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    st = ifb._abi.StageList(imf.build_stages(k1, 3) + imf.build_stages(k2, 3))
    device.imgradients(img_d.desc(), [ifb.DeviceArray.from_torch(gx).desc(), ifb.DeviceArray.from_torch(gy).desc()],
                       st, 3, ifb.Pad("reflect").to_abi(3))
    torch.cuda.synchronize()
    assert device.last_path() in ("stream2d_grad", "fused2d_grad")
    himg = ifb.n0f8(np.asfortranarray(host.transpose(2, 1, 0)))
    o1 = ifb.imfilter(himg, k1, "reflect", _library=oracle)
    o2 = ifb.imfilter(himg, k2, "reflect", _library=oracle)
    assert np.array_equal(gx.cpu().numpy().transpose(2, 1, 0), o1)
    assert np.array_equal(gy.cpu().numpy().transpose(2, 1, 0), o2)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.uint8, np.int16, np.int32])
def test_extrema_parity(ifb, oracle, device, dt):
    rng = np.random.default_rng(43)
    for shape, window in (((101,), (7,)), ((64, 45), (7, 7)), ((33, 20), (4, 3)), ((20, 19, 6), (3, 5, 1)), ((9, 8, 7), (2, 2, 2))):
        if np.dtype(dt).kind == "f":
            img = np.asfortranarray(rng.random(shape).astype(dt))
        else:
            img = np.asfortranarray(rng.integers(0, 200, size=shape).astype(dt))
        a = ifb.mapwindow(ifb.extrema, img, window)
        b = ifb.mapwindow(ifb.extrema, img, window, _library=oracle)
        assert np.array_equal(a["min"], b["min"]) and np.array_equal(a["max"], b["max"]), (shape, window)
        if all(w % 2 == 1 for w in window):
            for f in (ifb.minimum, ifb.maximum):
                for border in ("replicate", "reflect", ifb.Fill(5), ifb.Inner()):
                    x = ifb.mapwindow(f, img, window, border=border)
                    y = ifb.mapwindow(f, img, window, border=border, _library=oracle)
                    px = x.parent if isinstance(x, ifb.OffsetArray) else x
                    py = y.parent if isinstance(y, ifb.OffsetArray) else y
                    assert np.array_equal(px, py), (shape, window, border)


@pytest.mark.parametrize("dt", [np.uint8, np.int16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("w", [17, 31, 63, 101, 18, 64])
def test_running_extrema_van_herk(ifb, oracle, device, dt, w):
    """Windows wider than the register kernel, and every eltype other than Float32, run the van Herk / Gil-Werman kernel
    (csrc/extrema_vh.cu): O(1) comparisons per element whatever the window.  Bit-exact against the oracle's truncated-window
    ground truth (src/mapwindow.jl:388-473 semantics), odd and even widths, 1-D / 2-D / 3-D (window on the slowest axis too),
    Fill and Inner borders, N0f8 images."""
    rng = np.random.default_rng(seed_of("vh", str(np.dtype(dt)), w))

    def data(shape):
        if np.dtype(dt).kind == "f":
            return np.asfortranarray(rng.random(shape).astype(dt))
        return np.asfortranarray(rng.integers(0, 200, size=shape).astype(dt))
    cases = [((300,), (w,)), ((257, 130), (w, 1)), ((140, 211), (1, w)), ((150, 120), (w, 5)), ((40, 37, 130), (3, 1, w))]
    for shape, window in cases:
        img = data(shape)
        a = ifb.mapwindow(ifb.extrema, img, window)
        assert device.last_path() == "extrema_vh", (device.last_path(), shape, window)
        b = ifb.mapwindow(ifb.extrema, img, window, _library=oracle)
        assert np.array_equal(a["min"], b["min"]) and np.array_equal(a["max"], b["max"]), (shape, window)
    if w % 2 == 1:
        img = data((150, 140))
        for f in (ifb.minimum, ifb.maximum):
            for border in ("replicate", ifb.Fill(7), ifb.Inner()):
                x = ifb.mapwindow(f, img, (w, 9), border=border)
                assert device.last_path() == "extrema_vh"
                y = ifb.mapwindow(f, img, (w, 9), border=border, _library=oracle)
                px = x.parent if isinstance(x, ifb.OffsetArray) else x
                py = y.parent if isinstance(y, ifb.OffsetArray) else y
                assert np.array_equal(px, py), (w, border)
    if dt == np.uint8:                                   # N0f8 images (the reference's default 8-bit eltype)
        img8 = ifb.n0f8(rng.integers(0, 256, size=(200, 90), dtype=np.uint8))
        a = ifb.mapwindow(ifb.extrema, img8, (w, 3))
        b = ifb.mapwindow(ifb.extrema, img8, (w, 3), _library=oracle)
        assert device.last_path() == "extrema_vh"
        assert np.array_equal(np.asarray(a["min"]), np.asarray(b["min"])) and np.array_equal(np.asarray(a["max"]), np.asarray(b["max"]))


@pytest.mark.parametrize("force", ["fused2d", "generic"])
def test_slower_paths_stay_bit_exact(ifb, oracle, device, force, monkeypatch):
    """The smem-tiled kernel and the per-stage path are the fallbacks of the streamed kernel: force them."""
    monkeypatch.setenv("B2F_FORCE_PATH", force)
    rng = np.random.default_rng(47)
    raw = rng.integers(0, 256, size=(150, 97, 2), dtype=np.uint8)
    img = ifb.n0f8(raw)
    for border in BORDERS + [ifb.Fill(0.5), ifb.Inner()]:
        for kf in (ifb.KernelFactors.gaussian((2, 1, 0)), ifb.KernelFactors.gaussian((5, 6, 0))):
            pa, pb = _both(ifb, oracle, img, kf, border)
            assert np.array_equal(pa, pb), (force, border)
            assert device.last_path() == force
    f = np.asfortranarray(rng.random((90, 120), dtype=np.float32))
    kf = ifb.KernelFactors.gaussian((3, 3))
    pa, pb = _both(ifb, oracle, np.float32, f, kf, "symmetric")
    assert np.max(np.abs(pa - pb)) <= _tol([k.data.parent for k in kf], f)
    ga = ifb.imgradients(img, ifb.KernelFactors.sobel, "reflect")
    gb = ifb.imgradients(img, ifb.KernelFactors.sobel, "reflect", _library=oracle)
    for a, b in zip(ga, gb):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("taps", [(1, 1), (2, 3), (3, 3), (4, 2), (5, 5), (7, 7), (8, 6), (9, 9), (13, 13), (16, 11), (17, 17), (3, 17)])
@pytest.mark.parametrize("combo", ["f32-f32", "u8-f64", "f64-f64", "f32-f64", "n0f8-f32"])
def test_stream2d_tap_sweep(ifb, oracle, device, taps, combo):
    """Every instantiation of the streamed kernel (exact hot sizes and run-time buckets), ragged strip edges,
    asymmetric tap offsets, all border styles."""
    src, dst = combo.split("-")
    rng = np.random.default_rng(seed_of((taps, combo)))
    lx, ly = taps
    kx = ifb.OffsetArray.with_first(rng.standard_normal(lx), (-(lx // 2) + (lx % 3 == 0),))
    ky = ifb.OffsetArray.with_first(rng.standard_normal(ly).reshape(1, ly), (0, -(ly // 3)))
    for shape in ((203, 131, 2), (64, 40), (7, 5)):
        if src == "f32":
            img = np.asfortranarray(rng.random(shape, dtype=np.float32))
        elif src == "f64":
            img = np.asfortranarray(rng.random(shape))
        else:
            raw = np.asfortranarray(rng.integers(0, 256, size=shape, dtype=np.uint8))
            img = ifb.n0f8(raw) if src == "n0f8" else raw
        T = np.float32 if dst == "f32" else np.float64
        kern = (kx, ky) if len(shape) == 2 else (
            ifb.OffsetArray.with_first(kx.parent.reshape(lx, 1, 1), (kx.first[0], 0, 0)),
            ifb.OffsetArray.with_first(ky.parent.reshape(1, ly, 1), (0, ky.first[1], 0)))
        for border in BORDERS + [ifb.Fill(0.5 if src != "u8" else 2), ifb.Inner()]:
            if isinstance(border, str) and border == "reflect" and min(shape[:2]) < 2:
                continue
            pa, pb = _both(ifb, oracle, T, img, kern, border)
            if pa.size == 0:
                continue
            if lx > 1 and ly > 1 and (max(lx, ly) <= 16 or (lx, ly) == (17, 17)):
                assert device.last_path() == "stream2d", (device.last_path(), taps)
            if dst == "f64":
                assert np.array_equal(pa, pb), (taps, combo, shape, border)
            else:
                tol = _tol([kx.parent, ky.parent], np.asarray(img))
                assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= tol, (taps, combo, shape, border)


@pytest.mark.parametrize("ksize", [(2, 2), (3, 5), (7, 4), (16, 9), (27, 27), (32, 13), (5, 40)])
def test_dense2d_parity(ifb, oracle, device, ksize):
    """K2: dense non-separable kernels (Kernel.LoG-like), exact in Float64, tolerance in Float32."""
    rng = np.random.default_rng(seed_of(ksize))
    kx, ky = ksize
    kern = ifb.OffsetArray.with_first(rng.standard_normal((kx, ky)), (-(kx // 2), -(ky // 3)))
    for shape, dt in (((150, 70), "f32"), ((64, 64, 2), "f64"), ((33, 90), "u8"), ((9, 7), "f32")):
        if dt == "f32":
            img = np.asfortranarray(rng.random(shape, dtype=np.float32))
        elif dt == "f64":
            img = np.asfortranarray(rng.random(shape))
        else:
            img = np.asfortranarray(rng.integers(0, 256, size=shape, dtype=np.uint8))
        k = kern if len(shape) == 2 else ifb.OffsetArray.with_first(kern.parent.reshape(kx, ky, 1), kern.first + (0,))
        for border in BORDERS + [ifb.Fill(1.0), ifb.Inner()]:
            pa, pb = _both(ifb, oracle, np.float64, img, (k,), border)
            if pa.size == 0:
                continue
            assert device.last_path() == "dense2d", device.last_path()
            assert np.array_equal(pa, pb), (ksize, shape, dt, border)
            pa, pb = _both(ifb, oracle, np.float32, img, (k,), border)
            tol = _tol([kern.parent], np.asarray(img))
            assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= tol, (ksize, shape, dt, border)


def test_log3_circular_config3_small(ifb, oracle, device):
    """BASELINE config 3 in miniature: Kernel.LoG(3) (27x27, rank 2 -> dense path) with Pad(:circular)."""
    rng = np.random.default_rng(3)
    img = np.asfortranarray(rng.random((200, 160), dtype=np.float32))
    k = ifb.Kernel.LoG(3)
    pa, pb = _both(ifb, oracle, np.float32, img, k, "circular")
    assert device.last_path() == "dense2d"
    assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= _tol([k.parent], img)
    pa, pb = _both(ifb, oracle, img, k, "circular")           # reference-typed Float64 result: bit-exact
    assert pa.dtype == np.float64 and np.array_equal(pa, pb)


@pytest.mark.parametrize("border", BORDERS + ["fill"])
def test_sepnd_parity(ifb, oracle, device, border):
    """N-d separable cascades as chained streamed passes (3-D gaussian = BASELINE config 5 in miniature,
    single-axis stages, 1-D arrays, 4-D arrays); Fill exercises the pushed-through fill value."""
    rng = np.random.default_rng(seed_of(border))
    b = ifb.Fill(0.7) if border == "fill" else border
    cases = [
        (np.asfortranarray(rng.random((70, 50, 40), dtype=np.float32)), ifb.KernelFactors.gaussian((4, 4, 4))),
        (np.asfortranarray(rng.random((33, 41, 29))), ifb.KernelFactors.gaussian((1, 2, 3))),
        (np.asfortranarray(rng.random((20, 19, 18))), tuple(reversed(ifb.KernelFactors.gaussian((1, 1, 2))))),   # z, y, x order
        (rng.random(500), (ifb.centered(rng.random(7)),)),
        (np.asfortranarray(rng.random((64, 37))), (ifb.OffsetArray.with_first(rng.random((1, 5)), (0, -3)),)),    # axis 1 only
        (np.asfortranarray(rng.random((64, 37))), (ifb.OffsetArray.with_first(rng.random((4, 1)), (-1, 0)),)),    # axis 0 only
        (np.asfortranarray(rng.random((12, 11, 10, 9))), ifb.KernelFactors.gaussian((1, 1, 1, 1))),
        (np.asfortranarray(rng.integers(0, 256, size=(40, 30, 20), dtype=np.uint8)), ifb.KernelFactors.sobel((True, True, True), 3)),
    ]
    for img, kern in cases:
        if border == "fill" and img.dtype == np.uint8:
            b = ifb.Fill(3)
        pa, pb = _both(ifb, oracle, np.float64, img, kern, b)
        assert device.last_path() == "sepnd", (device.last_path(), img.shape)
        assert np.array_equal(pa, pb), (img.shape, border)
        pa, pb = _both(ifb, oracle, np.float32, img, kern, b)
        taps = [k.data.parent if isinstance(k, ifb.ReshapedOneD) else k.parent for k in kern]
        assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= _tol(taps, np.asarray(img)), (img.shape, border)
    g = ifb.imgradients(cases[0][0], ifb.KernelFactors.sobel, b)      # 3-D gradients: three 3-stage cascades
    go = ifb.imgradients(cases[0][0], ifb.KernelFactors.sobel, b, _library=oracle)
    for a, o in zip(g, go):
        assert np.max(np.abs(a.astype(np.float64) - o.astype(np.float64))) <= 1e-5


@pytest.mark.parametrize("border", ["symmetric", "replicate", "reflect", "circular", "fill"])
@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_slab_form_matches_whole_volume(ifb, oracle, device, border, T):
    """b2f_imfilter_slab on ONE device: cut a volume into 3 slabs, hand every slab its neighbours' RAW boundary
    planes as halos (what the NCCL exchange / the peer mapping delivers) and compare with the oracle on the whole
    volume.  Float32 runs the fused stream3d kernel, Float64 the gathered per-stage slab path (bit-exact)."""
    import torch
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    rng = np.random.default_rng(53)
    X, Y, Z, h = 48, 40, 60, 8
    vol = np.asfortranarray(rng.random((X, Y, Z)).astype(T))
    kf = ifb.KernelFactors.gaussian((4, 4, 4)) if T == np.float32 else ifb.KernelFactors.gaussian((4.0, 4.0, 4.0))
    b = ifb.Fill(0.3) if border == "fill" else ifb.Pad(border)
    ref = ifb.imfilter(T, vol, kf, b, _library=oracle)
    t_vol = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0))).cuda()
    st = ifb._abi.StageList(imf.build_stages(kf, 3))
    out = torch.empty_like(t_vol)
    bounds = [0, 17, 41, Z]
    circ = border == "circular"
    for i in range(3):
        z0, z1 = bounds[i], bounds[i + 1]
        lo = h if (i > 0 or circ) else 0
        hi = h if (i < 2 or circ) else 0
        own = t_vol[z0:z1].contiguous()
        hlo = t_vol[[(z % Z) for z in range(z0 - lo, z0)]].contiguous() if lo else None
        hhi = t_vol[[(z % Z) for z in range(z1, z1 + hi)]].contiguous() if hi else None
        o = torch.empty_like(own)
        device.imfilter_slab(ifb.DeviceArray.from_torch(own).desc(), ifb.DeviceArray.from_torch(o).desc(), st,
                             b.to_abi(3), Z, z0, hlo.data_ptr() if lo else 0, lo, hhi.data_ptr() if hi else 0, hi)
        assert device.last_path() == ("stream3d_slab" if T == np.float32 else "slab")
        out[z0:z1] = o
    got = out.cpu().numpy().transpose(2, 1, 0)
    if T == np.float64:
        assert np.array_equal(got, ref)
    else:
        assert np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) <= 1e-5
    with pytest.raises(ifb.DimensionMismatch):     # halo smaller than the kernel needs
        own = t_vol[17:41].contiguous()
        hlo, hhi = t_vol[15:17].contiguous(), t_vol[41:43].contiguous()
        o = torch.empty_like(own)
        device.imfilter_slab(ifb.DeviceArray.from_torch(own).desc(), ifb.DeviceArray.from_torch(o).desc(), st,
                             b.to_abi(3), Z, 17, hlo.data_ptr(), 2, hhi.data_ptr(), 2)


@pytest.mark.parametrize("border", ["symmetric", "replicate", "reflect", "circular", "fill"])
@pytest.mark.parametrize("sigma", [4, 2, 1])
def test_slab_xy_form_matches_whole_volume(ifb, oracle, device, border, sigma):
    """b2f_imfilter_slab_xy on ONE device: three slabs, each handed its own and its neighbours' boundary planes ALREADY filtered
    along x and y (what the sharded driver exchanges instead of raw halos).  The result must equal the unsharded fused kernel
    bit for bit (the march only skips stages whose results it is given) and the oracle within the Float32 tolerance.  sigma
    4 / 2 / 1 = 17 / 9 / 5 taps = halos of 8 / 4 / 2 planes (even and odd numbers of skipped planes per ring slot)."""
    import torch
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    rng = np.random.default_rng(91 + sigma)
    X, Y, Z = 64, 96, 60
    vol = np.asfortranarray(rng.random((X, Y, Z)).astype(np.float32))
    kf = ifb.KernelFactors.gaussian((sigma, sigma, sigma))
    h = 2 * sigma
    b = ifb.Fill(0.0) if border == "fill" else ifb.Pad(border)
    ref = ifb.imfilter(np.float32, vol, kf, b, _library=oracle)
    t_vol = torch.from_numpy(np.ascontiguousarray(vol.transpose(2, 1, 0))).cuda()
    st = ifb._abi.StageList(imf.build_stages(kf, 3))
    st_xy = ifb._abi.StageList(imf.build_stages(kf[:2], 3))
    DA = ifb.DeviceArray
    whole, t_xy = torch.empty_like(t_vol), torch.empty_like(t_vol)
    device.imfilter(DA.from_torch(t_vol).desc(), DA.from_torch(whole).desc(), st, b.to_abi(3), None, 0)
    assert device.last_path() == "stream3d"
    device.imfilter(DA.from_torch(t_vol).desc(), DA.from_torch(t_xy).desc(), st_xy, b.to_abi(3), None, 0)
    out = torch.empty_like(t_vol)
    bounds = [0, 17, 41, Z]
    circ = border == "circular"
    for i in range(3):
        z0, z1 = bounds[i], bounds[i + 1]
        lo = h if (i > 0 or circ) else 0
        hi = h if (i < 2 or circ) else 0
        own = t_vol[z0:z1].contiguous()
        xy_lo = t_xy[[(z % Z) for z in range(z0 - lo, z0 + h)]].contiguous()
        xy_hi = t_xy[[(z % Z) for z in range(z1 - h, z1 + hi)]].contiguous()
        o = torch.empty_like(own)
        device.imfilter_slab_xy(DA.from_torch(own).desc(), DA.from_torch(o).desc(), st, b.to_abi(3), Z, z0,
                                xy_lo.data_ptr(), lo, h, xy_hi.data_ptr(), h, hi)
        assert device.last_path() == "stream3d_slab"
        out[z0:z1] = o
    torch.cuda.synchronize()
    assert torch.equal(out, whole)
    got = out.cpu().numpy().transpose(2, 1, 0)
    assert np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))) <= 1e-5


@pytest.mark.parametrize("border", BORDERS + ["fill"])
def test_stream3d_parity(ifb, oracle, device, border, monkeypatch):
    """Fused 3-D separable kernel (BASELINE config 5 in miniature): exact 17^3 / 9^3 / 5^3 / 3^3 instantiations and
    the run-time-count one, tiles with x/y edges, odd widths (scalar loads / stores), arrays thinner than the halo,
    asymmetric and even-length factors; the two-pass `sepnd` path must agree within the same tolerance."""
    rng = np.random.default_rng(sum(map(ord, border)))
    b = ifb.Fill(0.7) if border == "fill" else border
    g = ifb.KernelFactors.gaussian
    asym = (ifb.ReshapedOneD(3, 0, ifb.OffsetArray.with_first(rng.random(4).astype(np.float32), (-1,))),
            ifb.ReshapedOneD(3, 1, ifb.OffsetArray.with_first(rng.random(6).astype(np.float32), (-4,))),
            ifb.ReshapedOneD(3, 2, ifb.OffsetArray.with_first(rng.random(3).astype(np.float32), (0,))))
    cases = [
        ((70, 50, 40), g((4, 4, 4))),
        ((200, 97, 23), g((4, 4, 4))),
        ((129, 33, 19), g((2, 2, 2))),
        ((65, 70, 12), g((1, 1, 1))),
        ((131, 37, 9), ifb.KernelFactors.sobel((True, True, True), 2)),
        ((33, 41, 29), g((1, 2, 3))),
        ((64, 64, 64), g((3, 2, 4))),
        ((5, 4, 3), g((4, 4, 4))),
        ((77, 66, 30), asym),
        ((256, 80, 70), g((4, 4, 4))),
    ]
    for shape, kern in cases:
        if border == "reflect" and min(shape) < 2:
            continue
        img = np.asfortranarray(rng.random(shape, dtype=np.float32))
        pa, pb = _both(ifb, oracle, np.float32, img, kern, b)
        assert device.last_path() == "stream3d", (device.last_path(), shape)
        taps = [k.data.parent for k in kern]
        err = np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64)))
        assert err <= _tol(taps, img), (shape, border, err)
    monkeypatch.setenv("B2F_FORCE_PATH", "sepnd")
    img = np.asfortranarray(rng.random((70, 50, 40), dtype=np.float32))
    pa, pb = _both(ifb, oracle, np.float32, img, g((4, 4, 4)), b)
    assert device.last_path() == "sepnd"
    assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= _tol([k.data.parent for k in g((4, 4, 4))], img)


@pytest.mark.parametrize("border", BORDERS + ["fill"])
def test_stream3d_more_tiles_than_sms(ifb, oracle, device, border):
    """The MIXED schedule of the fused 3-D kernel — whole waves of tiles marching every plane plus the remaining tiles cut
    into z-chunks (`bid >= nfull`, `kch > 1`; csrc/stream3d.cu) — needs more tiles than SMs: 544 x 576 x 40 is 17 x 9 = 153
    tiles of 32 x 64 (148 whole marches + 5 tiles x 5 chunks of 8 planes on a B200).  This is the schedule the 1024^3
    benchmark runs; every voxel is compared with the oracle."""
    rng = np.random.default_rng(seed_of("s3-mixed", border))
    shape = (544, 576, 40)
    img = np.asfortranarray(rng.random(shape, dtype=np.float32))
    kern = ifb.KernelFactors.gaussian((4, 4, 4))
    b = ifb.Fill(0.25) if border == "fill" else border
    pa, pb = _both(ifb, oracle, np.float32, img, kern, b)
    assert device.last_path() == "stream3d"
    err = np.abs(pa.astype(np.float64) - pb.astype(np.float64))
    tol = _tol([k.data.parent for k in kern], img)
    assert err.max() <= tol, (border, float(err.max()), np.unravel_index(err.argmax(), err.shape))


def test_extrema_full_hd_batch(ifb, oracle, device):
    """BASELINE config 4 at its real image size: mapwindow(extrema|minimum|maximum, img, (7,7)) on a 1920 x 1080 x 4 batch,
    bit-exact against the oracle (src/mapwindow.jl:388-473 semantics: window truncated at the image ends)."""
    rng = np.random.default_rng(4)
    img = np.asfortranarray(rng.random((1920, 1080, 4), dtype=np.float32))
    mm = ifb.mapwindow(ifb.extrema, img, (7, 7, 1))
    assert device.last_path().startswith("extrema")
    mo = ifb.mapwindow(ifb.extrema, img, (7, 7, 1), _library=oracle)
    assert np.array_equal(mm["min"], mo["min"]) and np.array_equal(mm["max"], mo["max"])
    assert np.array_equal(ifb.mapwindow(ifb.minimum, img, (7, 7, 1)), mo["min"])
    assert np.array_equal(ifb.mapwindow(ifb.maximum, img, (7, 7, 1)), mo["max"])


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.uint8, np.int32, np.int64])
def test_findlocalextrema_parity(ifb, oracle, device, dt):
    """Strict-peak scan (src/extrema.jl:125-162): GPU index lists == oracle index lists (same order), for plateaus
    (integer data with many ties), NaNs, even windows, per-axis edge flags, 1-D .. 4-D."""
    rng = np.random.default_rng(int(np.dtype(dt).itemsize) * 7 + 1)
    cases = [((1000,), (3,), (True,)), ((257, 130), (3, 3), (True, False)), ((64, 50, 33), (3, 3, 3), (False, True, True)),
             ((40, 30, 20), (1, 5, 2), (True, True, False)), ((9, 8, 7, 6), (3, 3, 3, 3), (True, True, True, True)),
             ((300, 300), (7, 1), (False, False)), ((2, 3), (3, 3), (False, True)), ((1, 5), (3, 3), (True, True))]
    for shape, window, edges in cases:
        if np.dtype(dt).kind == "f":
            A = rng.random(shape).astype(dt)
            if A.size > 50:
                A.ravel()[rng.integers(0, A.size, 5)] = np.nan
        else:
            A = rng.integers(0, 6, size=shape).astype(dt)
        A = np.asfortranarray(A)
        for f in (ifb.findlocalmaxima, ifb.findlocalminima):
            device.reset_launch_count()
            got = f(A, window=window, edges=edges)
            assert device.launch_count() > 0 and device.last_path() == "localextrema"
            assert got == f(A, window=window, edges=edges, _library=oracle), (shape, window, edges, f.__name__)


def test_blob_log_parity(ifb, oracle, device):
    """blob_LoG on random blobs: same peaks and σ as the oracle, amplitudes bit-equal in Float64 and within the Float32
    tolerance for Float32 images; the σ-stack stays on the GPU (multi-σ LoG + scan + gather)."""
    rng = np.random.default_rng(11)
    yy, xx = np.meshgrid(np.arange(96), np.arange(128))
    img = np.zeros((128, 96))
    for cx, cy, s in ((30, 20, 2.0), (80, 60, 4.0), (100, 30, 3.0)):
        img += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img += 0.01 * rng.random(img.shape)
    for T in (np.float64, np.float32):
        A = np.asfortranarray(img.astype(T))
        got = ifb.blob_LoG(A, [2.0, 3.0, 4.0, 6.0], rthresh=0.05)
        ref = ifb.blob_LoG(A, [2.0, 3.0, 4.0, 6.0], rthresh=0.05, _library=oracle)
        if T == np.float64:
            assert got == ref and len(got) >= 3
        else:       # Float32 LoG stack: peaks may only differ where the amplitudes tie within the tolerance
            strong = lambda bl: sorted((b.location, b.σ) for b in bl if b.amplitude > 0.1)
            assert strong(got) == strong(ref) and len(strong(got)) >= 3
            amp = {b.location: b.amplitude for b in ref}
            assert all(abs(b.amplitude - amp[b.location]) <= 1e-5 for b in got if b.location in amp)


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_na_border_parity(ifb, oracle, device, T):
    """NA() border (src/imfilter.jl:282-318): separable path (no NA present: Fill(0) + normalize_dims), inseparable path
    (NaNs present, or a dense kernel: two Fill(0) passes and the division), host and device-resident arrays."""
    import torch
    rng = np.random.default_rng(21)
    img = np.asfortranarray(rng.random((150, 90)).astype(T))
    holes = img.copy(order="F")
    holes.ravel(order="K")[rng.integers(0, img.size, 200)] = np.nan
    sep = ifb.KernelFactors.gaussian((2, 1.5))
    dense = ifb.Kernel.DoG((1.5, 1.5)) if hasattr(ifb.Kernel, "DoG") else ifb.Kernel.LoG(1.5)
    vol = np.asfortranarray(rng.random((40, 30, 20)).astype(T))
    for A, kern in ((img, sep), (holes, sep), (img, dense), (holes, dense), (vol, ifb.KernelFactors.gaussian((1, 1, 1)))):
        got = ifb.imfilter(A, kern, ifb.NA())
        ref = ifb.imfilter(A, kern, ifb.NA(), _library=oracle)
        assert got.dtype == ref.dtype
        if got.dtype == np.float64:
            assert np.array_equal(got, ref, equal_nan=True)
        else:
            assert np.array_equal(np.isnan(got), np.isnan(ref))
            assert np.nanmax(np.abs(got.astype(np.float64) - ref.astype(np.float64))) <= 1e-5
    # device-resident in and out
    t_in = torch.from_numpy(np.ascontiguousarray(holes.T)).cuda()
    t_out = torch.empty(t_in.shape, dtype=torch.float64 if T == np.float64 else torch.float32, device="cuda")
    if T == np.float64:
        ifb.imfilter_(ifb.DeviceArray.from_torch(t_out), ifb.DeviceArray.from_torch(t_in), sep, ifb.NA())
        torch.cuda.synchronize()
        ref = ifb.imfilter(holes, sep, ifb.NA(), _library=oracle)
        assert np.array_equal(t_out.cpu().numpy().T, ref, equal_nan=True)


def test_color_image_parity(ifb, oracle, device):
    """RGB{N0f8} / RGB{Float32} images = a leading channel axis (src/imfilter.jl:1131-1154 eltype arithmetic): each
    channel equals the scalar result of that channel, bit-exact in Float64."""
    rng = np.random.default_rng(31)
    raw = np.asfortranarray(rng.integers(0, 256, size=(3, 97, 61), dtype=np.uint8))
    for kern, border in ((ifb.KernelFactors.gaussian((2, 2)), "reflect"), (ifb.Kernel.LoG(1.0), "circular"),
                         (ifb.KernelFactors.sobel((True, True), 1), ifb.Fill(0)), (ifb.KernelFactors.gaussian((1, 2)), ifb.NA())):
        got = ifb.imfilter(ifb.ColorArray(raw), kern, border)
        ref = ifb.imfilter(ifb.ColorArray(raw), kern, border, _library=oracle)
        assert got.data.dtype == np.float64 and np.array_equal(got.data, ref.data, equal_nan=True)
        for c in range(3):
            one = ifb.imfilter(ifb.n0f8(np.asfortranarray(raw[c])), kern, border)
            assert np.array_equal(got.channel(c), one, equal_nan=True), (c, border)
    f = ifb.ColorArray(np.asfortranarray(rng.random((3, 64, 50), dtype=np.float32)))
    got = ifb.imfilter(np.float32, f, ifb.KernelFactors.gaussian((2, 2)), "symmetric")
    ref = ifb.imfilter(np.float32, f, ifb.KernelFactors.gaussian((2, 2)), "symmetric", _library=oracle)
    assert np.max(np.abs(got.data.astype(np.float64) - ref.data)) <= 1e-5


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8, np.int64])
def test_median_window_parity(ifb, oracle, device, dt):
    """mapwindow(median!, ...) (src/mapwindow.jl:270-333 + Statistics.median!): bit-equal to the oracle for odd and even
    windows, asymmetric ranges, every border (the Pad styles pad the window's in-image part), NaNs, 1-D .. 3-D."""
    rng = np.random.default_rng(int(np.dtype(dt).itemsize) + 40)
    cases = [((200,), (5,)), ((64, 50), (3, 3)), ((64, 50), (range(-2, 2), range(0, 3))), ((33, 20, 11), (3, 1, 5)),
             ((40, 30), (7, 7)), ((5, 4), (7, 9))]
    borders = ["replicate", "circular", "symmetric", "reflect", ifb.Fill(2), ifb.Inner()]
    for shape, window in cases:
        if np.dtype(dt).kind == "f":
            A = rng.random(shape).astype(dt)
            if A.size > 100:
                A.ravel()[rng.integers(0, A.size, 4)] = np.nan
        else:
            A = rng.integers(0, 50, size=shape).astype(dt)
        A = np.asfortranarray(A)
        for border in borders:
            if isinstance(border, ifb.Inner) and shape == (5, 4):
                continue
            device.reset_launch_count()
            got = ifb.mapwindow(ifb.median, A, window, border)
            assert device.launch_count() > 0 and device.last_path() == "median"
            ref = ifb.mapwindow(ifb.median, A, window, border, _library=oracle)
            g = got.parent if isinstance(got, ifb.OffsetArray) else got
            r = ref.parent if isinstance(ref, ifb.OffsetArray) else ref
            assert g.dtype == r.dtype == (np.float32 if dt == np.float32 else np.float64)
            assert np.array_equal(g, r, equal_nan=True), (shape, window, border)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.uint8, np.int64])
def test_window_reductions_parity(ifb, oracle, device, dt):
    """mapwindow(mean | sum | minimum | maximum | median!, ...; border, indices=strided ranges) — the generic window loop of the
    reference (src/mapwindow.jl:270-306) — bit-equal to the oracle: sums run in the window's memory order with the
    reference's accumulator type on both sides."""
    rng = np.random.default_rng(seed_of("winreduce", str(np.dtype(dt))))
    if np.dtype(dt).kind == "f":
        img = np.asfortranarray(rng.standard_normal((61, 47)).astype(dt))
        vol = np.asfortranarray(rng.standard_normal((20, 17, 9)).astype(dt))
    else:
        img = np.asfortranarray(rng.integers(0, 200, size=(61, 47)).astype(dt))
        vol = np.asfortranarray(rng.integers(0, 200, size=(20, 17, 9)).astype(dt))
    for f in (ifb.mean, ifb.sum_, ifb.minimum, ifb.maximum, ifb.median):
        for border in ("replicate", "reflect", "circular", "symmetric", ifb.Fill(3), ifb.Inner()):
            a = ifb.mapwindow(f, img, (5, 3), border=border)
            if f in (ifb.mean, ifb.sum_):
                assert device.last_path() == "winreduce"
            b = ifb.mapwindow(f, img, (5, 3), border=border, _library=oracle)
            pa = a.parent if isinstance(a, ifb.OffsetArray) else a
            pb = b.parent if isinstance(b, ifb.OffsetArray) else b
            assert pa.dtype == pb.dtype and np.array_equal(pa, pb, equal_nan=True), (f, border)
        inds = (range(2, 60, 3), range(1, 47, 2))
        a = ifb.mapwindow(f, img, (range(-2, 2), range(-1, 2)), indices=inds)
        b = ifb.mapwindow(f, img, (range(-2, 2), range(-1, 2)), indices=inds, _library=oracle)
        assert a.shape == (20, 23) and np.array_equal(a, b), f
        a = ifb.mapwindow(f, vol, (3, 1, 5), border="symmetric", indices=(range(1, 21, 4), range(1, 18), range(3, 9, 2)))
        b = ifb.mapwindow(f, vol, (3, 1, 5), border="symmetric", indices=(range(1, 21, 4), range(1, 18), range(3, 9, 2)), _library=oracle)
        assert np.array_equal(a, b), f


@pytest.mark.parametrize("border", BORDERS + ["fill"])
@pytest.mark.parametrize("dt", ["f32", "f64", "n0f8"])
def test_long_separable_factors(ifb, oracle, device, border, dt):
    """18 .. 256 taps per axis (csrc/longtap.cu): the reference's benchmark kernel KernelFactors.gaussian(sigma = 10) is 41 taps
    (benchmark/benchmarks.jl:45-49).  Float64 outputs bit-exact, Float32 within the stated tolerance; 2-D and 3-D, tap counts
    around the chunk size of 8, mixed with short factors, pads larger than the array."""
    rng = np.random.default_rng(seed_of(("long", border, dt)))
    g = ifb.KernelFactors.gaussian
    b = ifb.Fill(0.25) if border == "fill" else border
    cases = [((150, 131), g((10, 10))),                  # 41 x 41
             ((300, 40), g((5, 8))),                     # 21 x 33
             ((33, 70, 45), g((1, 6, 5))),               # 5, 25, 21: short x factor, long y and z
             ((20, 19, 50), g((7, 1, 16))),              # 29 taps on an axis of 20, 65 on an axis of 50
             ((260, 37), g((0, 9)))]                     # y only, 37 taps
    for shape, kern in cases:
        if dt == "n0f8":
            img = ifb.n0f8(rng.integers(0, 256, size=shape, dtype=np.uint8))
        else:
            img = np.asfortranarray(rng.random(shape).astype(np.float32 if dt == "f32" else np.float64))
        T = np.float32 if dt == "f32" else np.float64
        pa, pb = _both(ifb, oracle, T, img, kern, b)
        assert device.last_path() == "sepnd", (device.last_path(), shape)
        if T == np.float64:
            assert np.array_equal(pa, pb), (shape, border, dt)
        else:
            tol = _tol([k.data.parent for k in kern], img)
            assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= tol, (shape, border)
    # tap counts 18 .. 34 and 255 / 256 with an asymmetric (non-centred) factor along each axis in turn
    img = np.asfortranarray(rng.random((70, 61, 9)))
    for L in (18, 23, 24, 25, 26, 33, 34, 255, 256):
        taps = rng.random(L) - 0.3
        for axis in (0, 1, 2):
            k = ifb.ReshapedOneD(3, axis, ifb.OffsetArray.with_first(taps, (-(L // 3),)))
            pa, pb = _both(ifb, oracle, np.float64, img, (k,), b)
            assert device.last_path() == "sepnd"
            assert np.array_equal(pa, pb), (L, axis, border)


def test_accum_mode_fma(ifb, oracle, device):
    """b2f_set_accum_mode(B2F_ACCUM_FMA): Float64 outputs may come from fused multiply-adds.  Non-dyadic taps: within one
    rounding per tap of the oracle (and not all bit-equal: the fused kernel really ran); dyadic taps on N0f8 data (Sobel): the
    products are exact either way, so the result stays bit-equal.  The mode is per thread and restored by the context manager."""
    rng = np.random.default_rng(77)
    img = np.asfortranarray(rng.random((300, 211, 3), dtype=np.float32))
    for sig, L in ((3, 13), (1, 5), (4, 17)):
        kern = ifb.KernelFactors.gaussian((sig, sig, 0))
        exact, ref = _both(ifb, oracle, np.float64, img, kern, "replicate")
        assert np.array_equal(exact, ref)
        with ifb.accum_mode("fma"):
            fused = ifb.imfilter(np.float64, img, kern, "replicate")
            assert device.last_path() == "stream2d"
        tol = 1e-15 * float(np.prod([np.abs(k.data.parent).sum() for k in kern[:2]]))
        assert np.max(np.abs(fused - ref)) <= tol
        assert not np.array_equal(fused, ref), "the fused kernel did not run"
        again = ifb.imfilter(np.float64, img, kern, "replicate")
        assert np.array_equal(again, ref), "accum mode was not restored"
    raw = ifb.n0f8(rng.integers(0, 256, size=(257, 130), dtype=np.uint8))
    ref = ifb.imgradients(raw, ifb.KernelFactors.sobel, "reflect", _library=oracle)
    with ifb.accum_mode("fma"):
        g = ifb.imgradients(raw, ifb.KernelFactors.sobel, "reflect")
    for a, b in zip(g, ref):
        assert np.array_equal(a, b)
    assert device.set_accum_mode(0) == 0


@pytest.mark.parametrize("border", BORDERS + ["fill", "inner"])
def test_dense3d_parity(ifb, oracle, device, border):
    """One dense 3-D stage (csrc/dense3d.cu; reference loop src/imfilter.jl:624-669; its benchmark kernels 3x3x3 and the
    13x13x13 DoG, benchmark/benchmarks.jl:36-49): Float64 outputs bit-exact (taps in column-major order, separate multiply and
    add), Float32 within tolerance; asymmetric offsets, a 4-th batch axis, N0f8 input."""
    rng = np.random.default_rng(seed_of(("dense3d", border)))
    b = {"fill": ifb.Fill(0.3), "inner": ifb.Inner()}.get(border, border)
    k3 = ifb.centered(rng.random((3, 3, 3)) - 0.4)
    kdog = ifb.Kernel.DoG((2, 2, 2))
    kasym = ifb.OffsetArray.with_first(rng.random((5, 2, 3)) - 0.5, (-1, 0, -2))
    for shape, kern in (((40, 37, 21), k3), ((50, 33, 30), kdog), ((37, 20, 9), kasym), ((33, 9, 6, 3), k3)):
        img = np.asfortranarray(rng.random(shape))
        kk = kern
        if len(shape) == 4:
            kk = ifb.OffsetArray.with_first(kern.parent.reshape(kern.parent.shape + (1,)), tuple(kern.first) + (0,))
        pa, pb = _both(ifb, oracle, img, (kk,), b)
        assert device.last_path() == "dense3d", device.last_path()
        assert pa.dtype == np.float64 and np.array_equal(pa, pb), (shape, border)
    img32 = np.asfortranarray(rng.random((45, 31, 17), dtype=np.float32))
    pa, pb = _both(ifb, oracle, np.float32, img32, (kdog,), b)
    assert device.last_path() == "dense3d"
    assert np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))) <= _tol([kdog.parent], img32)
    raw = ifb.n0f8(np.asfortranarray(rng.integers(0, 256, size=(35, 18, 11), dtype=np.uint8)))
    pa, pb = _both(ifb, oracle, raw, (k3,), b)
    assert device.last_path() == "dense3d" and np.array_equal(pa, pb)


def test_pipelined_host_call_matches_plain_call(ifb, oracle, device):
    """Large host-to-host calls on PINNED arrays run as an upload / kernel / download pipeline over chunks of planes
    (csrc/api.cu, imfilter_host_pipelined).  Same result as the plain call on ordinary numpy arrays (bit for bit: the same
    kernels, chunked through the slab form), and as the oracle on sampled blocks; 3-D cascade with halos between the chunks,
    and a batch of 2-D images (no halo)."""
    import ctypes as C
    rng = np.random.default_rng(123)

    def pinned(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        device.check(device.dll.b2f_host_alloc(C.byref(p), n))
        ctype = {4: C.c_float, 8: C.c_double}[np.dtype(dtype).itemsize]
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(int(np.prod(shape)),))
        return p, a.reshape(shape, order="F")

    for shape, kern, border, T in (((256, 192, 330), ifb.KernelFactors.gaussian((4, 4, 4)), "symmetric", np.float32),
                                   ((200, 160, 520), ifb.KernelFactors.gaussian((2, 3, 1)), ifb.Fill(0.5), np.float64),
                                   ((512, 384, 140), ifb.KernelFactors.gaussian((3, 3, 0)), "replicate", np.float32)):
        pi, img = pinned(shape, np.float32)
        po, out = pinned(shape, T)
        try:
            img[...] = rng.random(shape, dtype=np.float32)
            ifb.imfilter_(out, img, kern, border)
            path = device.last_path()
            assert path in ("stream3d_slab", "slab"), path
            plain = ifb.imfilter(T, np.asfortranarray(np.array(img)), kern, border)
            assert device.last_path() not in ("stream3d_slab", "slab")
            if T == np.float64:
                assert np.array_equal(out, plain), shape
            else:
                tol = _tol([k.data.parent for k in kern], img)
                assert np.max(np.abs(out - plain)) <= tol, shape
            # the oracle on a block that spans a chunk boundary region and both faces of the last axis
            for z0, z1 in ((0, 12), (shape[2] // 2 - 6, shape[2] // 2 + 6), (shape[2] - 12, shape[2])):
                lo, hi = max(0, z0 - 8), min(shape[2], z1 + 8)
                blk = np.asfortranarray(np.array(img[:64, :48, lo:hi]))
                ref = ifb.imfilter(np.float64, blk, kern, border, _library=oracle)
                # compare away from the block's artificial x / y / z faces
                zz0 = z0 - lo if lo > 0 else 0
                zz1 = (z1 - lo) if hi < shape[2] else blk.shape[2]
                got = np.array(out[:40, :28, lo + zz0:lo + zz1], dtype=np.float64)
                want = ref[:40, :28, zz0:zz1]
                tol = (_tol([k.data.parent for k in kern], img) if T == np.float32 else 1e-12)
                assert np.max(np.abs(got - want)) <= tol, (shape, z0)
        finally:
            device.dll.b2f_host_free(pi)
            device.dll.b2f_host_free(po)
