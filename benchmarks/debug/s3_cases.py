"""Debug helper: run the stream3d parity cases one by one (GPU vs oracle) and print the error of each."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import imagefiltering_jl_b200 as ifb
from importlib import import_module
lib = import_module("imagefiltering_jl_b200._lib").lib()
oracle = ifb._abi.Library(os.path.join(ROOT, "oracle", "libb2f_oracle.so"))
g = ifb.KernelFactors.gaussian
rng = np.random.default_rng(1)
shapes = [((70, 50, 40), g((4, 4, 4))), ((200, 97, 23), g((4, 4, 4))), ((129, 33, 19), g((2, 2, 2))), ((65, 70, 12), g((1, 1, 1))),
          ((33, 41, 29), g((1, 2, 3))), ((64, 64, 64), g((3, 2, 4))), ((5, 4, 3), g((4, 4, 4))), ((256, 80, 70), g((4, 4, 4))),
          ((256, 160, 40), g((4, 4, 4))), ((128, 128, 50), g((2, 2, 2)))]
only = sys.argv[1:]
if os.environ.get('S3_CASE'):
    shapes = [shapes[int(i)] for i in os.environ['S3_CASE'].split(',')]
for border in (only or ["replicate", "symmetric", "circular", "reflect", "fill0", "fill"]):
    b = ifb.Fill(0.7) if border == "fill" else ifb.Fill(0) if border == "fill0" else border
    for shape, kern in shapes:
        img = np.asfortranarray(rng.random(shape, dtype=np.float32))
        try:
            pa = ifb.imfilter(np.float32, img, kern, b)
            pb = ifb.imfilter(np.float32, img, kern, b, _library=oracle)
            err = np.max(np.abs(pa.astype(np.float64) - pb.astype(np.float64)))
            print(border, shape, lib.last_path(), "err=%.3g" % err, "OK" if err < 1e-5 else "BAD", flush=True)
        except Exception as e:
            print(border, shape, "EXC", str(e)[:200], flush=True)
            sys.exit(1)
