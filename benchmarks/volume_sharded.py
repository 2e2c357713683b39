#!/usr/bin/env python
"""BASELINE config 5 across GPUs: 3-D gaussian((4,4,4)) on a Float32 volume with Pad(:symmetric), slab-sharded along
the last axis (strong scaling: the volume is fixed, every rank owns planes/P of it).

    python benchmarks/volume_sharded.py [--n 1024] [--steps 10] [--mode p2p|sendrecv]           # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/volume_sharded.py --mode p2p

One step = one sharded filter pass of the whole volume: entry barrier (the neighbours' inputs must be complete),
halo transport (peer loads fused into the kernel for "p2p", NCCL send/recv for "sendrecv") and the fused stream3d
kernel.  Timed with CUDA events per rank between barriers, max over ranks.  Before timing, a seam check filters a
small volume both sharded and whole (single GPU path, itself oracle-checked in tests/) and compares every plane.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import imagefiltering_jl_b200 as ifb
    from importlib import import_module
    from bench import peaks
    sh = import_module("imagefiltering_jl_b200.sharded")
    imf = import_module("imagefiltering_jl_b200.imfilter")
    lib = import_module("imagefiltering_jl_b200._lib").lib()
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--planes", type=int, default=0, help="planes along the sharded axis (default: n)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", default="p2p", choices=["p2p", "staged", "sendrecv"])
    ap.add_argument("--allreduce-barrier", action="store_true", help="p2p/staged: all-reduce entry barrier instead of the neighbour hand-shake")
    ap.add_argument("--nosync", action="store_true", help="p2p: skip the per-step entry barrier (static inputs)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kern = ifb.KernelFactors.gaussian((4, 4, 4))
    border = ifb.Pad("symmetric")

    # ---- seam check: sharded == whole on a small volume (all ranks hold the same seeded whole volume) ----------------
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    zs = max(16, 24) * world
    small = torch.rand((zs, 96, 160), device=dev, generator=g)
    whole = torch.empty_like(small)
    st = ifb._abi.StageList(imf.build_stages(kern, 3))
    lib.imfilter(ifb.DeviceArray.from_torch(small).desc(), ifb.DeviceArray.from_torch(whole).desc(), st, border.to_abi(3), None,
                 torch.cuda.current_stream().cuda_stream)
    first, cnt = sh.slab_bounds(zs, world, rank)
    f = sh.ShardedImfilter(small[first:first + cnt].contiguous(), kern, border, mode=args.mode)
    got = f.run()
    torch.cuda.synchronize()
    seam_err = float((got - whole[first:first + cnt]).abs().max())
    f.close()
    t = torch.tensor([seam_err], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    seam_err = float(t[0])
    assert seam_err < 1e-5, f"sharded result differs from the whole-volume result: {seam_err}"

    # ---- the timed volume ---------------------------------------------------------------------------------------------
    n = args.n
    nz = args.planes or n
    first, cnt = sh.slab_bounds(nz, world, rank)
    g.manual_seed(1000 + rank)
    slab = torch.rand((cnt, n, n), device=dev, generator=g)
    f = sh.ShardedImfilter(slab, kern, border, mode=args.mode, handshake=not args.allreduce_barrier)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        f.run(sync=not args.nosync)
    barrier()
    path = lib.last_path()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        f.run(sync=not args.nosync)
    b.record(stream)
    barrier()
    ms = a.elapsed_time(b) / args.steps
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    f.close()
    if rank == 0:
        hbm, which = peaks()
        npx = n * n * nz
        gbs = npx * 8 / (ms * 1e-3) / 1e9
        halo_mb = 2 * f.h_lo * n * n * 4 / 1e6 if world > 1 else 0.0
        print(json.dumps({
            "workload": "c5-sharded", "desc": f"{n}x{n}x{nz} f32 gaussian((4,4,4)) Pad(:symmetric), {world} slab(s) along the last axis",
            "n_gpus": world, "mode": args.mode, "entry_barrier": not args.nosync, "ms": ms,
            "gpixel_per_s": npx / (ms * 1e-3) / 1e9, "achieved_gbs_total": gbs, "hbm_frac_per_gpu": gbs / world / hbm,
            "peak_source": "of " + which, "halo_mb_per_gpu_per_direction": halo_mb / 2, "path": path,
            "seam_check_max_abs_err": seam_err, "scaling": "strong"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
