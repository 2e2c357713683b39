"""Worker of the slab-sharded tests (spawned once per rank; gloo control plane on 127.0.0.1).

Every rank builds the same seeded whole array, keeps its slab, runs the sharded filter and compares its planes
with the oracle's result on the WHOLE array (tests/ may use the oracle as the checker).  `use_device` selects the
product library on cuda:0 (both ranks share the one GPU; CUDA IPC works between processes on one device) or, for
the CPU suite, the oracle library standing in for the per-slab compute so that the partitioning, neighbour
selection, message order and halo bookkeeping of sharded.py are what is under test."""
import os
import sys
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def cases(ifb):
    g = ifb.KernelFactors.gaussian
    rng = np.random.default_rng(11)
    out = []
    for border in ("replicate", "circular", "symmetric", "reflect", ifb.Fill(0.4)):
        out.append(("f32-3d-%s" % (border,), np.float32, (40, 22, 37), g((2, 2, 2)), border, None))
    out.append(("f64-3d-asym", np.float64, (24, 18, 21), (g((1, 1, 1))[0], g((1, 1, 1))[1],
                ifb.ReshapedOneD(3, 2, ifb.OffsetArray.with_first(rng.random(4), (-1,)))), "symmetric", None))
    out.append(("f32-3d-17taps", np.float32, (70, 45, 40), g((4, 4, 4)), "symmetric", None))
    out.append(("f32-3d-tma", np.float32, (64, 80, 37), g((2, 2, 2)), "reflect", None))        # wide enough for the TMA / staged path
    out.append(("f32-3d-tma17", np.float32, (128, 96, 48), g((4, 4, 4)), "circular", None))
    out.append(("f32-3d-tma17-sym", np.float32, (64, 96, 40), g((4, 4, 4)), "symmetric", None))   # xy-filtered exchange at the faces
    out.append(("f32-3d-tma5-fill0", np.float32, (64, 88, 30), g((1, 1, 1)), ifb.Fill(0.0), None))
    # one-sided factors along the sharded axis: rank r needs planes of r+1 only (or of r-1 only); the neighbour relation and the
    # hand-shake stay symmetric, transfers of zero planes are skipped
    up = ifb.ReshapedOneD(3, 2, ifb.OffsetArray.with_first(rng.random(3), (0,)))
    down = ifb.ReshapedOneD(3, 2, ifb.OffsetArray.with_first(rng.random(5), (-4,)))
    out.append(("f64-3d-onesided-up", np.float64, (20, 12, 21), (g((1, 1, 0))[0], g((1, 1, 0))[1], up), "replicate", None))
    out.append(("f64-3d-onesided-down", np.float64, (20, 12, 21), (g((1, 1, 0))[0], down), "symmetric", None))
    out.append(("f32-3d-tma-onesided", np.float32, (64, 80, 30), (g((1, 1, 0))[0], g((1, 1, 0))[1], up), "reflect", None))
    out.append(("f32-2d", np.float32, (33, 29), g((1, 2)), "reflect", None))
    out.append(("f32-3d-uneven", np.float32, (20, 12, 31), g((1, 1, 2)), "circular", [20, 11]))
    out.append(("f64-3d-xy-only", np.float64, (20, 12, 16), (g((1, 1, 0))[0], g((1, 1, 0))[1]), "replicate", None))
    return out


def run(rank, world, port, use_device, modes, errq, env=None):
    try:
        os.environ.update(env or {})
        import torch
        import torch.distributed as dist
        import imagefiltering_jl_b200 as ifb
        from importlib import import_module
        sh = import_module("imagefiltering_jl_b200.sharded")
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        oracle = ifb._abi.Library(os.path.join(ROOT, "oracle", "libb2f_oracle.so"))
        lib = None if use_device else oracle
        if use_device:
            torch.cuda.set_device(0)
        failures = []
        for name, T, shape, kern, border, counts in cases(ifb):
            if "onesided" in name and os.environ.get("B2F_SHARD_XY") == "1":
                continue                        # the xy-filtered exchange takes two-sided cascades (csrc/sharded.cu); nothing new to test
            rng = np.random.default_rng(sum(map(ord, name)))
            whole = np.asfortranarray(rng.random(shape).astype(T))          # Julia order (X, Y, Z)
            ref = ifb.imfilter(T, whole, kern, border, _library=oracle)
            nz = shape[-1]
            if counts is not None and world == len(counts):
                first, n = sum(counts[:rank]), counts[rank]
            else:
                first, n = sh.slab_bounds(nz, world, rank)
            t_whole = torch.from_numpy(np.ascontiguousarray(whole.transpose()))   # (Z, Y, X), C order
            slab = t_whole[first:first + n].contiguous()
            if use_device:
                slab = slab.cuda()
            for mode in modes:
                f = sh.ShardedImfilter(slab, kern, border, mode=mode, _library=lib)
                assert (f.first, f.global_planes) == (first, nz), (f.first, f.global_planes, first, nz)
                got = f.run()
                if use_device:
                    torch.cuda.synchronize()
                got = got.cpu().numpy().transpose()
                f.close()
                want = ref[..., first:first + n]
                if T == np.float64:
                    ok = np.array_equal(got, want)
                else:
                    taps = [np.abs(np.asarray(k.data.parent, dtype=np.float64)).sum() for k in kern]
                    tol = 1e-5 * float(np.prod(taps)) * float(np.abs(whole).max())
                    ok = got.shape == want.shape and float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) <= tol
                    if not use_device:      # the oracle computes slabs exactly like whole arrays
                        ok = ok and np.array_equal(got, want)
                if not ok:
                    failures.append((name, mode, rank))
        dist.barrier()
        dist.destroy_process_group()
        errq.put((rank, failures))
    except Exception:
        errq.put((rank, ["EXC: " + traceback.format_exc()]))


def run_mapwindow(rank, world, port, use_device, modes, errq, env=None):
    """ShardedMapwindow: every rank's planes against mapwindow of the oracle on the whole array."""
    try:
        os.environ.update(env or {})
        import torch
        import torch.distributed as dist
        import imagefiltering_jl_b200 as ifb
        from importlib import import_module
        sh = import_module("imagefiltering_jl_b200.sharded")
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        oracle = ifb._abi.Library(os.path.join(ROOT, "oracle", "libb2f_oracle.so"))
        lib = None if use_device else oracle
        if use_device:
            torch.cuda.set_device(0)
        failures = []
        cases = [("ext-f32-3d", ifb.extrema, np.float32, (40, 22, 37), (5, 3, 7), "replicate"),
                 ("ext-u8-3d-wide", ifb.extrema, np.uint8, (33, 20, 64), (3, 5, 21), "replicate"),
                 ("max-i32-2d", ifb.maximum, np.int32, (50, 41), (3, 9), "symmetric"),
                 ("min-f64-3d-circ", ifb.minimum, np.float64, (20, 12, 31), (1, 3, 5), "circular"),
                 ("max-f32-3d-fill", ifb.maximum, np.float32, (24, 18, 30), (3, 3, 5), ifb.Fill(0.5)),
                 ("min-f32-3d-asym", ifb.minimum, np.float32, (24, 18, 30), (range(0, 1), range(-1, 2), range(-1, 4)), "reflect"),
                 ("mean-f32-3d", ifb.mean, np.float32, (18, 14, 25), (3, 3, 5), "replicate"),
                 ("sum-u8-3d", sum, np.uint8, (18, 14, 25), (3, 1, 3), "symmetric"),
                 ("median-f64-3d", ifb.median, np.float64, (12, 10, 22), (3, 3, 3), "replicate"),
                 ("ext-f32-3d-noz", ifb.extrema, np.float32, (30, 20, 12), (7, 7, 1), "replicate"),
                 ("max-f64-3d-onesided", ifb.maximum, np.float64, (14, 9, 23), (range(0, 1), range(-1, 2), range(0, 3)), "replicate"),
                 ("min-i32-3d-onesided", ifb.minimum, np.int32, (14, 9, 23), (range(0, 1), range(0, 1), range(-4, 1)), "symmetric")]
        for name, f, T, shape, window, border in cases:
            if "onesided" in name and use_device:
                continue                        # what one-sided windows change is host logic (neighbours, message matching): CPU suite
            rng = np.random.default_rng(sum(map(ord, name)))
            whole = np.asfortranarray((rng.random(shape) * 200).astype(T))
            ref = ifb.mapwindow(f, whole, window, border=border, _library=oracle)
            first, n = sh.slab_bounds(shape[-1], world, rank)
            t_whole = torch.from_numpy(np.ascontiguousarray(whole.transpose()))
            slab = t_whole[first:first + n].contiguous()
            if use_device:
                slab = slab.cuda()
            m = sh.ShardedMapwindow(f, slab, window, border, _library=lib)
            assert (m.first, m.global_planes) == (first, shape[-1])
            got = m.run()
            if use_device:
                torch.cuda.synchronize()
            if f is ifb.extrema:
                want = (ref["min"][..., first:first + n], ref["max"][..., first:first + n])
                ok = all(np.array_equal(g.cpu().numpy().transpose(), w) for g, w in zip(got, want))
            else:
                g = got.cpu().numpy().transpose()
                w = np.asarray(ref)[..., first:first + n]
                ok = g.dtype == w.dtype and np.array_equal(g, w)
            if not ok:
                failures.append((name, rank))
        dist.barrier()
        dist.destroy_process_group()
        errq.put((rank, failures))
    except Exception:
        errq.put((rank, ["EXC: " + traceback.format_exc()]))


def launch(world, use_device, modes, env=None, target=None):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target or run, args=(r, world, port, use_device, modes, q, env)) for r in range(world)]
    for p in procs:
        p.start()
    results = []
    try:
        for _ in procs:
            results.append(q.get(timeout=180))
    except Exception:
        for p in procs:
            if p.is_alive():
                p.kill()
        raise
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.kill()
    return results
