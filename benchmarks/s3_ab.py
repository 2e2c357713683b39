#!/usr/bin/env python
"""A/B timing of the fused 3-D kernel (BASELINE config 5) under the debugging knobs of csrc/stream3d.cu, one sub-process per
variant so that each reads its own environment.  Not a bench number: used to pick kernel variants on the same box.

    python benchmarks/s3_ab.py [planes] ["B2F_S3_V=1" "B2F_S3_V=2 B2F_S3_CS=0" ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(planes):
    import torch
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    vol = torch.rand((planes, 1024, 1024), device=dev, generator=g)
    out = torch.empty_like(vol)
    st = ifb._abi.StageList(imf.build_stages(ifb.KernelFactors.gaussian((4, 4, 4)), 3))
    bname = os.environ.get("AB_BORDER", "symmetric")
    b = (ifb.Fill(0.0) if bname == "fill" else ifb.Pad(bname)).to_abi(3)
    di, do = ifb.DeviceArray.from_torch(vol).desc(), ifb.DeviceArray.from_torch(out).desc()
    s = torch.cuda.current_stream()
    fn = lambda: lib.imfilter(di, do, st, b, None, s.cuda_stream)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for rep in range(3):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(5):
            fn()
        e.record(s)
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 5
        best = min(best, ms)
        tot += ms
    chk = float(out[planes // 2, 500, 300:308].double().sum())
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B2F_") or k.startswith("AB_")}, "planes": planes,
                      "ms_best": best, "ms_mean": tot / 3, "hbm_frac_best": planes * 1024 * 1024 * 8 / (best * 1e-3) / 6546.9e9,
                      "path": lib.last_path(), "checksum": chk}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        planes = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
        variants = sys.argv[2:] or ["B2F_S3_V=1", "B2F_S3_V=2"]
        for v in variants:
            env = dict(os.environ)
            for kv in v.split():
                k, val = kv.split("=")
                env[k] = val
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(planes)], env=env, check=False)
