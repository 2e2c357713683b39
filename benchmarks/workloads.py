#!/usr/bin/env python
"""Device-resident timings of every BASELINE.json config on ONE GPU (parity is tests/' job; this only times).

    python benchmarks/workloads.py [--steps K] [--only c1,c4]

Prints one JSON object per workload: ms per launch set, Gpixel/s, algorithmic GB/s and the fraction of the
measured HBM peak.  Inputs are larger than L2 (or the batch is) so no flush is needed; CUDA-event timed.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    from bench import ClockSampler, peaks
    imf = import_module("imagefiltering_jl_b200.imfilter")
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--clocks", action="store_true", help="also sample SM clocks / throttle reasons under each workload")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    hbm, which = peaks()
    DA = ifb.DeviceArray

    clk = {}

    def timeit(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        if args.clocks:      # a second, longer untimed loop under the clock sampler (~0.6 s of load)
            sm = ClockSampler(0)
            sm.start()
            import time
            t0 = time.time()
            while time.time() - t0 < 0.6:
                for _ in range(5):
                    fn()
                torch.cuda.synchronize()
            clk.update(sm.stop())
        return ms

    def report(name, desc, ms, npx, bytes_px, extra=None):
        gbs = npx * bytes_px / (ms * 1e-3) / 1e9
        d = {"workload": name, "desc": desc, "ms": ms, "gpixel_per_s": npx / (ms * 1e-3) / 1e9,
             "algorithmic_bytes_per_px": bytes_px, "achieved_gbs": gbs, "hbm_frac": gbs / hbm,
             "peak_source": "of " + which, "path": lib.last_path()}
        if extra:
            d.update(extra)
        if clk:
            d["clocks"] = dict(clk)
        print(json.dumps(d), flush=True)

    g = torch.Generator(device=dev)
    g.manual_seed(1)

    def filt(img_t, out_t, kern, border, n0f8=False):
        nd = img_t.dim()
        st = ifb._abi.StageList(imf.build_stages(kern, nd))
        di, do = DA.from_torch(img_t, n0f8=n0f8).desc(), DA.from_torch(out_t).desc()
        b = border.to_abi(nd)
        return lambda: lib.imfilter(di, do, st, b, None, sptr)

    if not only or "c1" in only:   # 2048^2 f32, gaussian((3,3)), Pad(:replicate); batch of 64 (2.1 GB of traffic)
        B = 64
        img = torch.rand((B, 2048, 2048), device=dev, generator=g)
        kf = ifb.KernelFactors.gaussian((3, 3, 0))
        for T, bpp in ((torch.float32, 8), (torch.float64, 12)):
            out = torch.empty((B, 2048, 2048), dtype=T, device=dev)
            ms = timeit(filt(img, out, kf, ifb.Pad("replicate")), args.steps)
            report("c1", f"64x 2048^2 f32, gaussian((3,3)) 13+13 taps, Pad(:replicate) -> {str(T)[6:]}", ms, B * 2048 * 2048, bpp)
        one_in, one_out = img[:1].contiguous(), torch.empty((1, 2048, 2048), device=dev)
        ms = timeit(filt(one_in, one_out, kf, ifb.Pad("replicate")), 50)
        report("c1-single", "one 2048^2 f32 image (fits L2; launch-bound)", ms, 2048 * 2048, 8)
        del img, out

    if not only or "c2" in only:   # Sobel gradients, 4096^2 N0f8, Pad(:reflect)
        B = 8
        img = torch.randint(0, 256, (B, 4096, 4096), dtype=torch.uint8, device=dev, generator=g)
        k1 = imf.build_stages(ifb.KernelFactors.sobel((True, True, False), 1), 3)
        k2 = imf.build_stages(ifb.KernelFactors.sobel((True, True, False), 2), 3)
        st = ifb._abi.StageList(k1 + k2)
        for T, bpp in ((torch.float64, 17), (torch.float32, 9)):
            o1 = torch.empty((B, 4096, 4096), dtype=T, device=dev)
            o2 = torch.empty_like(o1)
            di = DA.from_torch(img, n0f8=True).desc()
            do = [DA.from_torch(o1).desc(), DA.from_torch(o2).desc()]
            b = ifb.Pad("reflect").to_abi(3)
            ms = timeit(lambda: lib.imgradients(di, do, st, 3, b, sptr), args.steps)
            report("c2", f"8x 4096^2 N0f8 Sobel imgradients, Pad(:reflect) -> 2x {str(T)[6:]}", ms, B * 4096 * 4096, bpp)
            del o1, o2
        del img

    if not only or "c3" in only:   # dense 27x27 LoG on 8192^2 f32, Pad(:circular)
        img = torch.rand((8192, 8192), device=dev, generator=g)
        out = torch.empty_like(img)
        kern = (ifb.Kernel.LoG(3),)
        ms = timeit(filt(img, out, kern, ifb.Pad("circular")), max(2, args.steps // 5))
        fma = 8192 * 8192 * 729
        report("c3", "8192^2 f32, Kernel.LoG(3) 27x27 dense, Pad(:circular) -> f32", ms, 8192 * 8192, 8,
               {"gfma_per_s": fma / (ms * 1e-3) / 1e9, "fp32_pipe_frac": fma / (ms * 1e-3) / (148 * 128 * 1.965e9)})
        del img, out

    if not only or "c4" in only:   # mapwindow extrema / min / max 7x7 over 256 1920x1080 f32 images
        B = 256
        img = torch.rand((B, 1080, 1920), device=dev, generator=g)
        di = DA.from_torch(img).desc()
        wlo, whi = (-3, -3, 0), (3, 3, 0)
        b = ifb.Pad("replicate").to_abi(3)
        pair = torch.empty((B, 1080, 1920, 2), device=dev)
        dp = ifb._abi.make_array(pair.data_ptr(), ifb._abi.F32, (1920, 1080, B), (1, 1, 1), ifb._abi.DEVICE)
        ms = timeit(lambda: lib.mapwindow_extrema(di, dp, None, True, wlo, whi, b, sptr), args.steps)
        report("c4-extrema", "256x 1920x1080 f32, mapwindow(extrema, (7,7)) -> (min,max) tuples", ms, B * 1080 * 1920, 12)
        del pair
        o = torch.empty((B, 1080, 1920), device=dev)
        do = DA.from_torch(o).desc()
        ms = timeit(lambda: lib.mapwindow_extrema(di, do, None, False, wlo, whi, b, sptr), args.steps)
        report("c4-min", "256x 1920x1080 f32, mapwindow(minimum, (7,7))", ms, B * 1080 * 1920, 8)
        ms = timeit(lambda: lib.mapwindow_extrema(di, None, do, False, wlo, whi, b, sptr), args.steps)
        report("c4-max", "256x 1920x1080 f32, mapwindow(maximum, (7,7))", ms, B * 1080 * 1920, 8)
        del img, o

    if not only or "c5" in only:   # 3-D gaussian((4,4,4)) on a 1024^3 f32 volume, Pad(:symmetric)
        n = 1024
        vol = torch.rand((n, n, n), device=dev, generator=g)
        out = torch.empty_like(vol)
        kf = ifb.KernelFactors.gaussian((4, 4, 4))
        ms = timeit(filt(vol, out, kf, ifb.Pad("symmetric")), max(2, args.steps // 2))
        report("c5", "1024^3 f32, gaussian((4,4,4)) 17x3 taps, Pad(:symmetric) -> f32", ms, n ** 3, 8)

    if not only or "f1" in only:   # strict-peak scan and the blob_LoG pipeline (synchronous calls: wall clock)
        import time
        img = torch.rand((8192, 8192), device=dev, generator=g)
        A = DA.from_torch(img)
        for _ in range(2):
            pk = ifb.findlocalmaxima(A, as_array=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            pk = ifb.findlocalmaxima(A, as_array=True)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        report("f1-peaks", f"findlocalmaxima, 8192^2 f32, 3x3 window -> {len(pk)} peaks (flags + scan + ordered scatter + D2H of the list)",
               ms, 8192 * 8192, 5)
        del img, A
        blob = torch.rand((2048, 2048), device=dev, generator=g)
        B = DA.from_torch(blob)
        for _ in range(2):
            bl = ifb.blob_LoG(B, [1.0, 2.0, 3.0], rthresh=0.5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            bl = ifb.blob_LoG(B, [1.0, 2.0, 3.0], rthresh=0.5)
        ms = (time.perf_counter() - t0) / 3 * 1e3
        report("f1-blob", f"blob_LoG, 2048^2 f32, 3 sigmas (9x9, 17x17, 27x27 dense LoG + scan) -> {len(bl)} blobs", ms, 2048 * 2048, 4 + 3 * 4)


if __name__ == "__main__":
    main()
