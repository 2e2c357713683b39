#!/usr/bin/env python
"""Runs ONE workload a few times (for ncu captures; never a bench number).  python benchmarks/prof_one.py c5 [planes]"""
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
    which = sys.argv[1] if len(sys.argv) > 1 else "c5"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 160
    dev = torch.device("cuda", 0)
    DA = ifb.DeviceArray
    if which == "c5":
        vol = torch.rand((n, 1024, 1024), device=dev)
        out = torch.empty_like(vol)
        st = ifb._abi.StageList(imf.build_stages(ifb.KernelFactors.gaussian((4, 4, 4)), 3))
        b = ifb.Pad("symmetric").to_abi(3)
        fn = lambda: lib.imfilter(DA.from_torch(vol).desc(), DA.from_torch(out).desc(), st, b, None, 0)
    elif which == "c1":
        img = torch.rand((n, 2048, 2048), device=dev)
        out = torch.empty_like(img)
        st = ifb._abi.StageList(imf.build_stages(ifb.KernelFactors.gaussian((3, 3, 0)), 3))
        b = ifb.Pad("replicate").to_abi(3)
        fn = lambda: lib.imfilter(DA.from_torch(img).desc(), DA.from_torch(out).desc(), st, b, None, 0)
    elif which == "c3":
        img = torch.rand((8192, 8192), device=dev)
        out = torch.empty_like(img)
        st = ifb._abi.StageList(imf.build_stages((ifb.Kernel.LoG(3),), 2))
        b = ifb.Pad("circular").to_abi(2)
        fn = lambda: lib.imfilter(DA.from_torch(img).desc(), DA.from_torch(out).desc(), st, b, None, 0)
    elif which == "c4":
        B = n if len(sys.argv) > 2 else 64
        img = torch.rand((B, 1080, 1920), device=dev)
        pair = torch.empty((B, 1080, 1920, 2), device=dev)
        dp = ifb._abi.make_array(pair.data_ptr(), ifb._abi.F32, (1920, 1080, B), (1, 1, 1), ifb._abi.DEVICE)
        b = ifb.Pad("replicate").to_abi(3)
        fn = lambda: lib.mapwindow_extrema(DA.from_torch(img).desc(), dp, None, True, (-3, -3, 0), (3, 3, 0), b, 0)
    elif which == "c1f64":
        img = torch.rand((n, 2048, 2048), device=dev)
        out = torch.empty((n, 2048, 2048), device=dev, dtype=torch.float64)
        st = ifb._abi.StageList(imf.build_stages(ifb.KernelFactors.gaussian((3, 3, 0)), 3))
        b = ifb.Pad("replicate").to_abi(3)
        lib.set_accum_mode(int(os.environ.get("PROF_ACCUM", "0")))
        fn = lambda: lib.imfilter(DA.from_torch(img).desc(), DA.from_torch(out).desc(), st, b, None, 0)
    else:
        raise SystemExit("unknown workload")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(lib.last_path())


if __name__ == "__main__":
    main()
