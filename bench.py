#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 FIR path (BASELINE.json metric: imfilter Gpixel/s and
% of the HBM roofline, next to the CPU reference path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): imgradients(img, KernelFactors.sobel, Pad(:reflect)) on
4096x4096 N0f8 images, reference-typed outputs (two Float64 gradient planes, bit-exact mode).
One step = one fused launch over a batch of BATCH images per GPU (weak scaling: every rank owns
its own batch; the path has no exchange step, so there is no collective in the data path).

  value      Gpixel/s, inputs and outputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D + launch + D2H per step
  roofline   algorithmic bytes (17 B/px = 1 in + 2x8 out) / average launch time vs measured HBM peak
  cpu_baseline  the CPU oracle's CPUThreads(FIRTiled) restatement on the host cores (bounded sample)

`--impl reference` times that CPU restatement alone (Julia is not in the image).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W = H = 4096
BATCH = 8            # images per GPU per step, device-resident arm (in 134 MB, out 2.1 GB: both > L2)
E2E_BATCH = 2        # images per GPU per step, host-buffer arm
BYTES_PER_PX = 1 + 2 * 8
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            v = d.get("hbm_gbs") or d.get("hbm_gb_s") or d.get("hbm")
            if isinstance(v, dict):
                v = v.get("value") or v.get("gbs")
            if v:
                return float(v), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def sobel_stages(ifb, ndim=3):
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    ext = (True, True) + (False,) * (ndim - 2)
    k1 = ifb.KernelFactors.sobel(ext, 1)
    k2 = ifb.KernelFactors.sobel(ext, 2)
    return ifb._abi.StageList(imf.build_stages(k1, ndim) + imf.build_stages(k2, ndim)), ndim


# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, threads=None):
    """CPUThreads(Algorithm.FIRTiled()) restatement (oracle/oracle.cpp b2f_oracle_imfilter_tiled) on ONE
    4096x4096 N0f8 image per step: both Sobel gradients, each padded and filtered independently, exactly
    as imgradients does (reference src/specialty.jl:47-51)."""
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    imf = import_module("imagefiltering_jl_b200.imfilter")
    path = os.path.join(ROOT, "oracle", "libb2f_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ifb._abi.Library(path)
    fn = lib.dll.b2f_oracle_imfilter_tiled
    fn.argtypes = [C.POINTER(ifb._abi.b2f_array), C.POINTER(ifb._abi.b2f_array), C.POINTER(ifb._abi.b2f_stage),
                   C.c_int32, C.POINTER(ifb._abi.b2f_border), C.POINTER(C.c_int64), C.c_int32]
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(2)
    raw = np.asfortranarray(rng.integers(0, 256, size=(W, H), dtype=np.uint8))
    img = ifb._abi.numpy_array_desc(raw, (1, 1), ifb._abi.N0F8)
    outs = [np.empty((W, H), dtype=np.float64, order="F") for _ in range(2)]
    odesc = [ifb._abi.numpy_array_desc(o, (1, 1)) for o in outs]
    border = ifb.Pad("reflect").to_abi(2)
    stl = [ifb._abi.StageList(imf.build_stages(ifb.KernelFactors.sobel((True, True), d), 2)) for d in (1, 2)]
    tile = (C.c_int64 * 4)(64, 64, 1, 1)     # 32 KiB Float64 tiles (L1-sized, TiledIteration.padded_tilesize's intent)

    def step():
        for d in range(2):
            rc = fn(C.byref(img), C.byref(odesc[d]), stl[d].arr, 2, C.byref(border), tile, threads)
            if rc != 0:
                raise RuntimeError(lib.dll.b2f_last_error().decode())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return (W * H * steps) / dt / 1e9, dt / steps * 1e3, threads


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    gpx, ms, threads = cpu_reference_run(steps, max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "imfilter_gpixel_per_s", "value": gpx, "unit": "Gpixel/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "imgradients Sobel 4096x4096 N0f8 Pad(:reflect) -> 2x Float64 (BASELINE configs[1])",
                   "images_per_step": 1, "note": "CPU restatement of CPUThreads(Algorithm.FIRTiled()); Julia unavailable"},
        "cpu_baseline": {"value": gpx, "unit": "Gpixel/s", "cores": threads, "kind": "port",
                         "sample": "one 4096x4096 image per step (both gradients), %d steps" % steps},
        "e2e": {"value": gpx, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def ours_arm(args):
    import torch
    import torch.distributed as dist
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    lib = import_module("imagefiltering_jl_b200._lib").lib()   # raises if the CUDA extension is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    g = torch.Generator(device=dev)
    g.manual_seed(2 + rank)
    img = torch.randint(0, 256, (BATCH, H, W), dtype=torch.uint8, device=dev, generator=g)
    gx = torch.empty((BATCH, H, W), dtype=torch.float64, device=dev)
    gy = torch.empty_like(gx)
    stages, nd = sobel_stages(ifb, 3)
    border = ifb.Pad("reflect").to_abi(3)
    d_img = ifb.DeviceArray.from_torch(img, n0f8=True).desc()
    d_out = [ifb.DeviceArray.from_torch(gx).desc(), ifb.DeviceArray.from_torch(gy).desc()]

    def step():
        lib.imgradients(d_img, d_out, stages, 3, border, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from before the warm-up to the end of the timed region (the timed region of a
    # short run is only a few ms long: sampling it alone could return no sample at all)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    for _ in range(args.warmup):
        step()
    barrier()
    assert lib.last_path().endswith("_grad"), lib.last_path()
    kernel_path = lib.last_path()
    t_load = time.time()
    while time.time() - t_load < 0.3:       # untimed: keep the GPU under the same load until the sampler has readings
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    lib.reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record(stream)
    for a, b in ev:
        a.record(stream)
        step()
        b.record(stream)
    t_end.record(stream)
    barrier()
    launches = lib.launch_count()
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    nin = E2E_BATCH * H * W
    hptr = C.c_void_p()
    lib.check(lib.dll.b2f_host_alloc(C.byref(hptr), nin))
    hin = np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_uint8)), shape=(nin,))
    hin[:] = np.random.default_rng(100 + rank).integers(0, 256, size=nin, dtype=np.uint8)
    houts = []
    for _ in range(2):
        p = C.c_void_p()
        lib.check(lib.dll.b2f_host_alloc(C.byref(p), nin * 8))
        houts.append(p)
    h_img = ifb._abi.make_array(hptr.value, ifb._abi.N0F8, (W, H, E2E_BATCH), (1, 1, 1), ifb._abi.HOST)
    h_out = [ifb._abi.make_array(p.value, ifb._abi.F64, (W, H, E2E_BATCH), (1, 1, 1), ifb._abi.HOST) for p in houts]

    def e2e_step():
        lib.imgradients(h_img, h_out, stages, 3, border, sptr)   # H2D, launch, D2H, sync

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    check = float(np.ctypeslib.as_array(C.cast(houts[0], C.POINTER(C.c_double)), shape=(4,))[1])

    t = torch.tensor([total_ms, e2e_ms, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms = (float(x) for x in t.tolist())

    if rank == 0:
        hbm, which = peaks()
        px_step = BATCH * W * H
        value = world * px_step * args.steps / (total_ms * 1e-3) / 1e9
        achieved = px_step * BYTES_PER_PX / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("stream2d_grad_bench_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        try:
            gpx, ms, threads = cpu_reference_run(steps=3, warmup=1)
            cpu = {"value": gpx, "unit": "Gpixel/s", "cores": threads, "kind": "port",
                   "sample": "3 steps of one 4096x4096 N0f8 image (both Sobel gradients, Float64), "
                             "CPUThreads(FIRTiled) restatement, %.0f ms/step" % ms}
        except Exception as e:   # the baseline is a report, never a dependency of the GPU numbers
            cpu = {"value": None, "unit": "Gpixel/s", "cores": 0, "kind": "port", "sample": "failed: %s" % e}
        line = {
            "metric": "imfilter_gpixel_per_s", "value": value, "unit": "Gpixel/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "imgradients Sobel 4096x4096 N0f8 Pad(:reflect) -> 2x Float64 (BASELINE configs[1])",
                       "images_per_gpu_per_step": BATCH, "parallelism": "batch-sharded x%d, no collective" % world,
                       "l2": "working set per step 2.3 GB per GPU, larger than the 126 MB L2",
                       "kernel": kernel_path + " (one launch per step)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "peak_source": "of " + which,
                         "algorithmic_bytes_per_launch": px_step * BYTES_PER_PX, "launch_ms": kern_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": world * E2E_BATCH * W * H * e2e_steps / (e2e_ms * 1e-3) / 1e9, "unit": "Gpixel/s",
                    "h2d_bytes_per_step": E2E_BATCH * W * H, "d2h_bytes_per_step": E2E_BATCH * W * H * 16,
                    "images_per_gpu_per_step": E2E_BATCH, "steps": e2e_steps, "checksum": check},
            "gpu_launches": launches * world,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for p in [hptr] + houts:
        lib.dll.b2f_host_free(p)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours_arm(args)


if __name__ == "__main__":
    main()
