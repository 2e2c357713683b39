import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    oracle = ifb._abi.Library(os.path.join(ROOT, "oracle", "libb2f_oracle.so"))
    shape = tuple(int(x) for x in sys.argv[1].split("x"))
    sig = tuple(int(x) for x in sys.argv[2].split(","))
    border = sys.argv[3]
    img = np.asfortranarray(np.random.default_rng(1).random(shape, dtype=np.float32))
    kern = ifb.KernelFactors.gaussian(sig)
    pa = ifb.imfilter(np.float32, img, kern, border)
    pb = ifb.imfilter(np.float32, img, kern, border, _library=oracle)
    print("err=%.3g" % np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64))), lib.last_path())
else:
    for shape, sig in [("64x64x64", "4,4,4"), ("64x64x64", "3,2,4"), ("64x64x40", "3,2,4"), ("64x128x64", "3,2,4"), ("64x64x64", "4,2,4"),
                       ("64x64x64", "3,4,4"), ("64x64x64", "4,4,3"), ("128x128x64", "3,3,3"), ("64x96x64", "4,4,4"), ("64x96x64", "3,2,4")]:
        for border in ["replicate", "circular"]:
            r = subprocess.run([sys.executable, __file__, shape, sig, border], capture_output=True, text=True)
            print(shape, sig, border, (r.stdout.strip() or r.stderr.strip()[-150:]).replace("\n", " "), flush=True)
