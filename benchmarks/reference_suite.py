#!/usr/bin/env python
"""The reference's OWN benchmark suite (benchmark/benchmarks.jl) on one B200, case by case, with the reference's names.

For every case two timings (CUDA events / wall clock around synchronous calls, best of 5 rounds):
  device_ms   arrays resident in HBM, output preallocated (imfilter! form) — the kernel path alone;
  host_ms     the public call on ordinary (pageable) numpy arrays, result returned as a numpy array: H2D + kernels + D2H,
              what a drop-in user of `imfilter(img, kernel, "replicate", Algorithm.FIR())` sees.
Cases outside the accelerated path are listed with the reason (ROF, arbitrary window functions).  Prints one JSON line
per case.  python benchmarks/reference_suite.py [substring]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    KF, K = ifb.KernelFactors, ifb.Kernel
    rng = np.random.default_rng(0)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    f32 = np.float32

    def best_of(fn, sync, rounds=5, inner=5):
        for _ in range(2):
            fn()
        sync()
        best = 1e9
        for _ in range(rounds):
            t0 = time.perf_counter()
            for _ in range(inner):
                fn()
            sync()
            best = min(best, (time.perf_counter() - t0) / inner)
        return best * 1e3

    def emit(name, **kw):
        print(json.dumps({"case": name, **kw}), flush=True)

    kerninsep = {1: ifb.centered(np.array([-1.0, 0.0, 1.0])),
                 2: ifb.centered(np.array([[1 / 5, 1 / 4, 1 / 7], [1 / 2, 1 / 3, -1 / 11], [-1 / 25, 1 / 9, -1 / 7]])),
                 3: ifb.centered(rng.random((3, 3, 3)))}
    for sz in ((100, 100), (2048, 2048), (2048,), (100, 100, 100)):
        nd = len(sz)
        szs = "x".join(map(str, sz))
        imgs = {"F32": np.asfortranarray(rng.random(sz, dtype=f32)),
                "N0f8": ifb.n0f8(np.asfortranarray(rng.integers(0, 256, size=sz, dtype=np.uint8)))}
        trues, twos = (True,) * nd, (2,) * nd
        kerns = {"densesmall": (kerninsep[nd],), "denselarge": (K.DoG(twos),), "factoredsmall": KF.sobel(trues, 1),
                 "factoredlarge": KF.gaussian(tuple(f32(10.0) for _ in sz)),
                 "IIRGaussian": KF.IIRGaussian(tuple(f32(10.0) for _ in sz))}
        for aname, img in imgs.items():
            for kname, kern in kerns.items():
                name = f"{kname}_{aname}_{szs}"
                if only and only not in name:
                    continue
                try:
                    host_ms = best_of(lambda: ifb.imfilter(img, kern, "replicate"), lambda: None)
                    path = lib.last_path()
                    T = ifb.filter_type(img, kern)
                    raw = img.raw if hasattr(img, "raw") else img
                    t_in = torch.from_numpy(np.ascontiguousarray(np.asarray(raw).transpose())).cuda()
                    t_out = torch.empty(t_in.shape, dtype=torch.float32 if T == np.float32 else torch.float64, device="cuda")
                    d_in = ifb.DeviceArray.from_torch(t_in, n0f8=aname == "N0f8")
                    d_out = ifb.DeviceArray.from_torch(t_out)
                    dev_ms = best_of(lambda: ifb.imfilter_(d_out, d_in, kern, "replicate"), torch.cuda.synchronize, inner=20)
                    emit(name, path=path, out_eltype=str(np.dtype(T)), device_ms=dev_ms, host_ms=host_ms,
                         device_gpixel_per_s=t_in.numel() / dev_ms / 1e6)
                except Exception as e:                                  # a report: name the gap, keep going
                    emit(name, error=f"{type(e).__name__}: {e}")
            name = f"FFT_{aname}_{szs}"                          # imfilter(img, (Kernel.DoG(twos),), "replicate", Algorithm.FFT())
            if not only or only in name:
                try:
                    fft, kern = ifb.Algorithm.FFT(), (K.DoG(twos),)
                    host_ms = best_of(lambda: ifb.imfilter(img, kern, "replicate", fft), lambda: None)
                    raw = img.raw if hasattr(img, "raw") else img
                    t_in = torch.from_numpy(np.ascontiguousarray(np.asarray(raw).transpose())).cuda()
                    t_out = torch.empty(t_in.shape, dtype=torch.float64, device="cuda")
                    d_in = ifb.DeviceArray.from_torch(t_in, n0f8=aname == "N0f8")
                    d_out = ifb.DeviceArray.from_torch(t_out)
                    dev_ms = best_of(lambda: ifb.imfilter_(ifb.CUDALibs(fft), d_out, d_in, kern, "replicate"), torch.cuda.synchronize, inner=10)
                    emit(name, path=lib.last_path(), out_eltype="float64", device_ms=dev_ms, host_ms=host_ms,
                         device_gpixel_per_s=t_in.numel() / dev_ms / 1e6)
                except Exception as e:
                    emit(name, error=f"{type(e).__name__}: {e}")
    # mapwindow group (benchmark/benchmarks.jl:21-35)
    img1d, img2d, img3d = rng.standard_normal(1000), np.asfortranarray(rng.standard_normal((30, 30))), np.asfortranarray(rng.standard_normal((10, 11, 12)))
    for name, f, im, w in (("extrema", ifb.extrema, img2d, (5, 5)), ("maximum", ifb.maximum, img2d, (5, 5)), ("minimum", ifb.minimum, img2d, (5, 5)),
                           ("median!", ifb.median, img2d, (5, 5)), ("mean, small window", ifb.mean, img1d, (3,)),
                           ("mean, large window", ifb.mean, img3d, (5, 5, 5))):
        if only and only not in name:
            continue
        ms = best_of(lambda: ifb.mapwindow(f, im, w), lambda: None)
        emit("mapwindow/" + name, path=lib.last_path(), host_ms=ms)
    for name in ("mapwindow/cheap f, tiny window", "mapwindow/expensive f"):
        if not only or only in name:
            emit(name, skipped="arbitrary Julia window functions cannot cross the C ABI")
    if not only:
        emit("ROF/PrimalDual_*", skipped="ImageFiltering.Models (ROF) is outside the FIR / min-max hot path")


if __name__ == "__main__":
    main()
