import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import imagefiltering_jl_b200 as ifb
from importlib import import_module
lib = import_module("imagefiltering_jl_b200._lib").lib()
imf = import_module("imagefiltering_jl_b200.imfilter")
dev = torch.device("cuda", 0)
img = torch.rand((2048, 2048), device=dev)
A = ifb.DeviceArray.from_torch(img)
def T(name, fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter()-t0)/n*1e3:.3f} ms", flush=True)
    return r
out = torch.empty_like(img)
for s in (1.0, 2.0, 3.0):
    k = ifb.Kernel.LoG((s, s))
    st = ifb._abi.StageList(imf.build_stages(imf.factorkernel(k), 2))
    T(f"imfilter LoG({s}) path", lambda: lib.imfilter(A.desc(), ifb.DeviceArray.from_torch(out).desc(), st, ifb.Pad("reflect").to_abi(2)))
    print(lib.last_path())
stack = torch.rand((2048, 2048, 3), device=dev)
S = ifb.DeviceArray.from_torch(stack)
pk = T("findlocalextrema stack raw", lambda: lib.findlocalextrema(S.desc(), False, (3, 3, 3), (True, False, False)))
print(len(pk))
T("gather", lambda: lib.gather(S.desc(), pk))
T("maxabs", lambda: lib.maxabs(A.desc()))
T("malloc+free 50MB", lambda: lib.free(lib.malloc(50_000_000)))
T("blob_LoG total", lambda: ifb.blob_LoG(A, [1.0, 2.0, 3.0], rthresh=0.5))
big = torch.rand((8192, 8192), device=dev)
B = ifb.DeviceArray.from_torch(big)
pk = T("findlocalextrema 8192^2 raw", lambda: lib.findlocalextrema(B.desc(), False, (3, 3), (True, True)))
print(len(pk))
T("findlocalmaxima 8192^2 as_array", lambda: ifb.findlocalmaxima(B, as_array=True))
