#!/usr/bin/env python
"""Times the reference's `factoredlarge` benchmark kernel — KernelFactors.gaussian(sigma = 10), 41 taps per axis
(benchmark/benchmarks.jl:45-49) — and neighbours on one GPU: device-resident arrays, CUDA events, best of 5 x 10 launches.
Prints one JSON line per case with the per-pass roofline (one pass per stage: sizeof(in) + sizeof(out) bytes per element).

    python benchmarks/longtap_time.py [substring of the case names to run]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    dev = torch.device("cuda", 0)
    DA = ifb.DeviceArray
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6546.9))
    cases = [("factoredlarge_F32_2048x2048", (2048, 2048), (10, 10), torch.float32, None),
             ("factoredlarge_F32_100x100x100", (100, 100, 100), (10, 10, 10), torch.float32, None),
             ("factoredlarge_F64_2048x2048", (2048, 2048), (10, 10), torch.float64, None),
             ("gaussian10_F32_8192x8192", (8192, 8192), (10, 10), torch.float32, None),
             ("gaussian10_F32_512^3", (512, 512, 512), (10, 10, 10), torch.float32, None),
             ("gaussian6_F32_8192x8192_fused2d", (8192, 8192), (6, 6), torch.float32, None),
             ("gaussian6_F32_8192x8192_sepnd", (8192, 8192), (6, 6), torch.float32, "sepnd"),
             ("gaussian4_F32_8192x8192_stream2d", (8192, 8192), (4, 4), torch.float32, None)]
    if len(sys.argv) > 1:
        cases = [c for c in cases if sys.argv[1] in c[0]]
    for name, shape, sig, dt, force in cases:
        if force:
            os.environ["B2F_FORCE_PATH"] = force
        else:
            os.environ.pop("B2F_FORCE_PATH", None)
        img = torch.rand(tuple(reversed(shape)), device=dev, dtype=dt)
        out = torch.empty_like(img)
        kern = ifb.KernelFactors.gaussian(sig)
        st = ifb._abi.StageList(imf.build_stages(kern, len(shape)))
        b = ifb.Pad("replicate").to_abi(len(shape))
        di, do = DA.from_torch(img).desc(), DA.from_torch(out).desc()
        fn = lambda: lib.imfilter(di, do, st, b, None, 0)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        n = img.numel()
        esz = img.element_size()
        passes = len(shape)
        path = lib.last_path()
        if path in ("stream2d", "fused2d", "stream3d"):
            passes = 1
        bytes_ = n * 2 * esz * passes
        print(json.dumps({"case": name, "taps": [len(k.data.parent) for k in kern], "path": path, "ms": best,
                          "gpixel_per_s": n / (best * 1e-3) / 1e9, "passes": passes,
                          "hbm_frac_of_passes": bytes_ / (best * 1e-3) / 1e9 / hbm}), flush=True)
        del img, out


if __name__ == "__main__":
    main()
