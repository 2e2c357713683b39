#!/usr/bin/env python
"""Times the reference's IIRGaussian benchmark (benchmark/benchmarks.jl:44,51: KernelFactors.IIRGaussian(sigma = 10) on
100^2 / 2048^2 / 100^3 Float32 arrays, "replicate") on one GPU: device-resident arrays, CUDA events, best of 5 x 10 cascades.
One cascade = one b2f_iir launch per axis (the later ones in place).  4 array passes per axis (see csrc/iir.cu)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    dev = torch.device("cuda", 0)
    DA = ifb.DeviceArray
    hbm = 6546.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        hbm = float(json.load(open(p)).get("hbm_gbs", hbm))
    for shape in ((100, 100), (2048, 2048), (100, 100, 100), (8192, 8192), (512, 512, 512)):
        nd = len(shape)
        kern = ifb.KernelFactors.IIRGaussian(tuple(np.float32(10.0) for _ in shape))
        img = torch.rand(tuple(reversed(shape)), device=dev)
        out = torch.empty_like(img)
        di, do = DA.from_torch(img).desc(), DA.from_torch(out).desc()
        b = ifb.Pad("replicate").to_abi(nd)
        coefs = [(k.Npre, k.data.coefficients()) for k in kern]

        def fn():
            src = di
            for ax, c in coefs:
                lib.iir(src, do, ax, c, b)
                src = do
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
        print(json.dumps({"case": "IIRGaussian_F32_" + "x".join(map(str, shape)), "ms": best, "gpixel_per_s": n / (best * 1e-3) / 1e9,
                          "launches": nd, "hbm_frac_of_4_passes_per_axis": n * 4 * 4 * nd / (best * 1e-3) / 1e9 / hbm}), flush=True)
        del img, out


if __name__ == "__main__":
    main()
