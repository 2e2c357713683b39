"""Slab-sharded N-d separable `imfilter` across GPUs (BASELINE config 5).

The array's LAST Julia axis (the slowest one: whole planes are contiguous) is partitioned over the
ranks of a `torch.distributed` group, one process per GPU.  Rank p owns planes
[first_p, first_p + n_p).  A cascade whose stages reach `lo` planes below and `hi` planes above
(`accumulate_padding`, reference src/border.jl:614-642) needs that many RAW input planes of each
neighbour; everything else — the other axes' borders, the outer faces of the volume — is local.
The result equals the owned planes of `imfilter(CPU1(Algorithm.FIR()), whole_array, kernel, border)`.

Two halo transports, same kernel entry (`b2f_imfilter_slab`, include/b2f.h):

  "p2p"       the neighbours' slabs are mapped into this process once (CUDA IPC over NVLink/NVSwitch,
              `b2f_ipc_export/open`); the filter kernel reads the boundary planes of its neighbours
              directly with peer loads while it streams its own planes — the exchange is fused into
              the compute kernel, no halo buffers, no copy kernels.  Per call only a stream-ordered
              barrier is needed (the neighbours' inputs must be complete before they are read).
  "staged"    the neighbours' slabs are mapped as for "p2p", but the boundary planes are pulled into local halo buffers by
              the COPY ENGINES on a side stream (dense copies at link speed, no tile over-fetch) while the filter kernel
              is already marching; a CTA waits on a one-byte flag only when it is about to read its first halo plane
              (`b2f_imfilter_slab_staged`).  Same barrier as "p2p".
  "sendrecv"  `dist.batch_isend_irecv` of the boundary planes into receive buffers (NCCL over NVLink;
              gloo on CPU tensors in the tests), then the same kernel with the buffers as halos.

The data path has no other collective.  Tensors are C-contiguous torch tensors whose reversed
shape is the Julia shape: (Zloc, Y, X) holds the Julia array (X, Y, Zloc).
"""
from __future__ import annotations

import numpy as np

from . import _abi
from ._abi import ArgumentError, DimensionMismatch, NotSupportedError
from .border import Fill, Pad, borderinstance
from .imfilter import build_stages, factorkernel, kernel_extent

_TORCH_TO_DT = None


def _dt_of(t):
    global _TORCH_TO_DT
    import torch
    if _TORCH_TO_DT is None:
        _TORCH_TO_DT = {torch.uint8: _abi.U8, torch.float32: _abi.F32, torch.float64: _abi.F64}
    if t.dtype not in _TORCH_TO_DT:
        raise NotSupportedError(f"slab eltype {t.dtype} is not supported")
    return _TORCH_TO_DT[t.dtype]


def slab_bounds(n_planes: int, world: int, rank: int):
    """Balanced contiguous partition of `n_planes` planes: -> (first, count) of `rank`."""
    base, extra = divmod(n_planes, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def halo_extent(stages, ndim: int):
    """Planes needed below / above along the last axis (Pad{0}(kernel): src/border.jl:602-606)."""
    first, last = kernel_extent(stages, ndim)
    return max(0, -first[ndim - 1]), max(0, last[ndim - 1])


def _desc(t, n0f8=False):
    dt = _abi.N0F8 if n0f8 else _dt_of(t)
    dims = tuple(reversed(t.shape))
    return _abi.make_array(t.data_ptr(), dt, dims, (1,) * len(dims), _abi.DEVICE if t.is_cuda else _abi.HOST)


class ShardedImfilter:
    """Plan of one slab-sharded `imfilter` over fixed input/output slabs; `run()` executes it.

        f = ShardedImfilter(slab, KernelFactors.gaussian((4, 4, 4)), Pad("symmetric"), group=g)
        out = f.run()          # any number of times (the input slab may be refilled in between)
        f.close()
    """

    def __init__(self, slab, kernel, border="replicate", *, out=None, group=None, mode="auto",
                 out_dtype=None, n0f8=False, handshake=True, _library=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.handshake = handshake          # p2p / staged: neighbour flag words instead of an all-reduce as the entry barrier
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if _library is not None:
            self.lib = _library
        else:
            from ._lib import lib
            self.lib = lib()                      # raises when the CUDA extension is missing: no CPU path
        if not slab.is_contiguous():
            raise ValueError("slab must be contiguous")
        if slab.dim() < 2 or slab.dim() > _abi.MAXDIM:
            raise NotSupportedError("slab-sharded arrays need 2..4 dimensions")
        self.slab, self.n0f8 = slab, n0f8
        self.ndim = slab.dim()
        border = borderinstance(border)
        if not isinstance(border, (Pad, Fill)) or border.lo:
            raise NotSupportedError("the sharded path takes Pad(style) / Fill(value) borders derived from the kernel")
        self.border = border
        if not isinstance(kernel, tuple):
            kernel = factorkernel(kernel)
        self.stages_list = build_stages(kernel, self.ndim)
        self.stages = _abi.StageList(self.stages_list)
        self.h_lo, self.h_hi = halo_extent(self.stages_list, self.ndim)
        # who owns what: every rank's plane count
        n_mine = int(slab.shape[0])
        counts = [n_mine]
        if self.world > 1:
            counts = [None] * self.world
            dist.all_gather_object(counts, n_mine, group=group)
        self.counts = counts
        self.first = int(sum(counts[:self.rank]))
        self.global_planes = int(sum(counts))
        circ = isinstance(border, Pad) and border.style == "circular"
        self.lower = self.rank - 1 if self.rank > 0 else (self.world - 1 if circ and self.world > 1 else None)
        self.upper = self.rank + 1 if self.rank + 1 < self.world else (0 if circ and self.world > 1 else None)
        # The neighbour relation is SYMMETRIC: a rank hand-shakes with both adjacent ranks as soon as the cascade reaches along
        # the sharded axis in either direction (a one-sided kernel makes rank r need planes of r+1 while r+1 needs none of r's —
        # r+1 must still tell r that its input is complete).  Transfers of zero planes are skipped on both sides.
        if self.h_lo == 0 and self.h_hi == 0:
            self.lower = self.upper = None
        for nb, need in ((self.lower, self.h_lo), (self.upper, self.h_hi)):
            if nb is not None and counts[nb] < need:
                raise DimensionMismatch(f"rank {nb} owns {counts[nb]} planes but its neighbour needs a halo of {need}")
        if out is None:
            odt = out_dtype or (slab.dtype if slab.dtype != torch.uint8 else torch.float32)
            out = torch.empty(slab.shape, dtype=odt, device=slab.device)
        if tuple(out.shape) != tuple(slab.shape) or not out.is_contiguous():
            raise DimensionMismatch("out must be a contiguous tensor with the slab's shape")
        self.out = out
        self.plane_elems = int(np.prod(slab.shape[1:]))
        if mode == "auto":
            mode = "driver" if (slab.is_cuda and self.world > 1 and self.lib.is_device_library()) else "sendrecv"
        if mode not in ("driver", "p2p", "staged", "sendrecv"):
            raise ArgumentError(f"unknown halo transport {mode!r}")
        self.mode = mode
        self._opened = []
        self.halo_lo_ptr = self.halo_hi_ptr = 0
        self.recv_lo = self.recv_hi = None
        self._nccl = self.world > 1 and dist.get_backend(group) == "nccl"
        self._flag = torch.zeros(1, dtype=torch.int32, device=slab.device) if self._nccl else None
        self._epoch = 0
        self._ctx = None
        if self.world > 1 and mode == "driver":
            self._setup_driver()
        elif self.world > 1 and (self.lower is not None or self.upper is not None or mode in ("p2p", "staged")):
            if mode == "p2p":
                self._setup_p2p()
            elif mode == "staged":
                self._setup_p2p()
                self.peer_lo_ptr, self.peer_hi_ptr = self.halo_lo_ptr, self.halo_hi_ptr
                self._setup_sendrecv()                 # local halo buffers; halo_*_ptr now point at them
                self._flags = torch.zeros(4, dtype=torch.uint8, device=slab.device)    # lo early, lo rest, hi, (pad)
                self._side = torch.cuda.Stream(device=slab.device)
                self._ev_go, self._ev_done = torch.cuda.Event(), torch.cuda.Event()
            else:
                self._setup_sendrecv()

    # -- transports -------------------------------------------------------------------------------------
    def _setup_p2p(self):
        if not self.slab.is_cuda:
            raise ArgumentError('halo transport "p2p" needs CUDA tensors')
        handle, offset = self.lib.ipc_export(self.slab.data_ptr())
        # neighbour hand-shake words: [0] written by my lower neighbour, [1] by my upper neighbour (step counters)
        # (their own cudaMalloc: an IPC handle names a whole allocation, and torch packs small tensors into shared blocks)
        self._sync_ptr = self.lib.malloc(8)
        self.lib.memset_async(self._sync_ptr, 0, 8, 0)
        self.lib.check(self.lib.dll.b2f_sync())
        self._step = 0
        self._peer_sync = {}
        sh, so = self.lib.ipc_export(self._sync_ptr)
        infos = [None] * self.world
        self.dist.all_gather_object(infos, (handle, offset, sh, so), group=self.group)
        for nb in {self.lower, self.upper} - {None}:
            base = self.lib.ipc_open(infos[nb][2], infos[nb][3])
            self._opened.append((base, infos[nb][3]))
            self._peer_sync[nb] = base
        infos = [(h, o) for h, o, _, _ in infos]
        esz = self.slab.element_size()
        lower_base = None
        if self.lower is not None:      # the last h_lo planes of the lower neighbour lie just below mine
            h, off = infos[self.lower]
            base = lower_base = self.lib.ipc_open(h, off)
            self._opened.append((base, off))
            self.halo_lo_ptr = base + (self.counts[self.lower] - self.h_lo) * self.plane_elems * esz
        if self.upper is not None:      # the first h_hi planes of the upper neighbour lie just above mine
            h, off = infos[self.upper]
            if self.upper == self.lower and lower_base is not None:
                base = lower_base
            else:
                base = self.lib.ipc_open(h, off)
                self._opened.append((base, off))
            self.halo_hi_ptr = base

    def _setup_driver(self):
        """The C driver (b2f_shard_*, csrc/sharded.cu): this class only moves the 256-byte blobs between the ranks."""
        import ctypes as C
        if not self.slab.is_cuda:
            raise ArgumentError('halo transport "driver" needs CUDA tensors')
        ctx = C.c_void_p()
        self.lib.check(self.lib.dll.b2f_shard_ctx_create(C.byref(ctx), self.rank, self.world))
        self._ctx = ctx
        blob = C.create_string_buffer(256)
        self.lib.check(self.lib.dll.b2f_shard_ctx_export(ctx, C.c_void_p(self.slab.data_ptr()), int(self.slab.shape[0]), blob))
        blobs = [None] * self.world
        self.dist.all_gather_object(blobs, blob.raw, group=self.group)
        lo = C.create_string_buffer(blobs[self.lower], 256) if self.lower is not None else None
        hi = C.create_string_buffer(blobs[self.upper], 256) if self.upper is not None else None
        self.lib.check(self.lib.dll.b2f_shard_ctx_connect(ctx, lo, hi))

    def _setup_sendrecv(self):
        t = self.torch
        shp = tuple(self.slab.shape[1:])
        if self.lower is not None:
            self.recv_lo = t.empty((self.h_lo,) + shp, dtype=self.slab.dtype, device=self.slab.device)
            self.halo_lo_ptr = self.recv_lo.data_ptr()
        if self.upper is not None:
            self.recv_hi = t.empty((self.h_hi,) + shp, dtype=self.slab.dtype, device=self.slab.device)
            self.halo_hi_ptr = self.recv_hi.data_ptr()

    def _exchange(self):
        """Boundary planes -> neighbours' receive buffers.  Message order between a pair of ranks: the
        sender posts [to lower, to upper], the receiver [from upper, from lower] (a circular wrap with two
        ranks exchanges two messages between the same pair)."""
        d = self.dist
        ops = []
        n = self.slab.shape[0]
        # gloo moves host memory only: CUDA slabs are staged through host copies there (tests; NCCL sends in place)
        stage = self.slab.is_cuda and not self._nccl
        send_lo = self.slab[:self.h_hi] if (self.lower is not None and self.h_hi > 0) else None          # -> lower neighbour's hi halo
        send_hi = self.slab[n - self.h_lo:] if (self.upper is not None and self.h_lo > 0) else None      # -> upper neighbour's lo halo
        recv_hi = self.recv_hi if (self.recv_hi is not None and self.h_hi > 0) else None
        recv_lo = self.recv_lo if (self.recv_lo is not None and self.h_lo > 0) else None
        if stage:
            send_lo = send_lo.cpu() if send_lo is not None else None
            send_hi = send_hi.cpu() if send_hi is not None else None
            recv_hi = self.torch.empty(recv_hi.shape, dtype=recv_hi.dtype) if recv_hi is not None else None
            recv_lo = self.torch.empty(recv_lo.shape, dtype=recv_lo.dtype) if recv_lo is not None else None
        if send_lo is not None:
            ops.append(d.P2POp(d.isend, send_lo, self._grank(self.lower), group=self.group))
        if send_hi is not None:
            ops.append(d.P2POp(d.isend, send_hi, self._grank(self.upper), group=self.group))
        if recv_hi is not None:
            ops.append(d.P2POp(d.irecv, recv_hi, self._grank(self.upper), group=self.group))
        if recv_lo is not None:
            ops.append(d.P2POp(d.irecv, recv_lo, self._grank(self.lower), group=self.group))
        if ops:
            for w in d.batch_isend_irecv(ops):
                w.wait()
        if stage:
            if recv_hi is not None:
                self.recv_hi.copy_(recv_hi)
            if recv_lo is not None:
                self.recv_lo.copy_(recv_lo)

    def _grank(self, r):
        return self.dist.get_global_rank(self.group, r) if self.group is not None else r

    def barrier(self, full=False):
        """All ranks' inputs are complete / all ranks have finished reading: stream-ordered on NCCL (a
        4-byte all-reduce on the current stream), host-side otherwise."""
        if self.world == 1:
            return
        if self.mode == "driver" and self._ctx is not None:
            import ctypes as C
            stream = self.torch.cuda.current_stream().cuda_stream
            return self.lib.check(self.lib.dll.b2f_shard_handshake(self._ctx, C.c_void_p(stream)))
        if not full and self.mode in ("p2p", "staged") and getattr(self, "_peer_sync", None) is not None and self.handshake:
            return self._neighbour_handshake()
        if self._nccl:
            self.dist.all_reduce(self._flag, group=self.group)
        else:
            if self.slab.is_cuda:
                self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)

    def _neighbour_handshake(self):
        """Only the neighbours' data is read, so only they are synchronised: write my step counter into their flag words
        (a 32-bit stream-ordered write over NVLink) and make the stream wait until mine have reached it.  No kernel, no
        collective; before REFILLING a slab call `barrier(full=True)`."""
        self._step += 1
        stream = self.torch.cuda.current_stream().cuda_stream
        if self.lower is not None:
            self.lib.stream_write32(self._peer_sync[self.lower] + 4, self._step, stream)     # I am its upper neighbour
        if self.upper is not None:
            self.lib.stream_write32(self._peer_sync[self.upper], self._step, stream)         # I am its lower neighbour
        if self.lower is not None:
            self.lib.stream_wait_geq32(self._sync_ptr, self._step, stream)
        if self.upper is not None:
            self.lib.stream_wait_geq32(self._sync_ptr + 4, self._step, stream)

    # -- execution --------------------------------------------------------------------------------------
    def run(self, sync=True):
        """One sharded filter pass -> `self.out`.  `sync=False` skips the entry barrier of the p2p transport
        (only valid when the neighbours' inputs are known to be complete and unchanged)."""
        t = self.torch
        stream = t.cuda.current_stream().cuda_stream if self.slab.is_cuda else 0
        if self.world > 1 and self.mode == "driver":
            import ctypes as C
            img, out, b = _desc(self.slab, self.n0f8), _desc(self.out), self.border.to_abi(self.ndim)
            self.lib.check(self.lib.dll.b2f_imfilter_sharded(self._ctx, C.byref(img), C.byref(out), self.stages.arr, self.stages.n, C.byref(b),
                                                             self.global_planes, self.first, C.c_void_p(stream)))
            return self.out
        if self.world > 1 and self.mode == "staged":
            return self._run_staged(sync, stream)
        if self.world > 1:
            if self.mode == "p2p":
                if sync:
                    self.barrier()
            else:
                self._exchange()
        self.lib.imfilter_slab(
            _desc(self.slab, self.n0f8), _desc(self.out), self.stages, self.border.to_abi(self.ndim),
            self.global_planes, self.first,
            self.halo_lo_ptr, self.h_lo if self.lower is not None else 0,
            self.halo_hi_ptr, self.h_hi if self.upper is not None else 0, stream)
        return self.out

    def _run_staged(self, sync, stream):
        """barrier -> [side stream: peer -> local halo copies, each followed by its flag byte] || [main stream: the kernel,
        whose CTAs wait for a flag only when they reach a halo plane] -> main stream joins the side stream."""
        t = self.torch
        if sync:
            self.barrier()
        self._epoch = self._epoch % 255 + 1
        main = t.cuda.current_stream()
        self._ev_go.record(main)
        self._side.wait_event(self._ev_go)
        side = self._side.cuda_stream
        esz = self.slab.element_size()
        flags = self._flags.data_ptr()
        # copy order = the order of need: the rows of the lower halo that the first wave of tiles reads, then the upper halo
        # (read when the first marches end), then the rest of the lower halo
        row_bytes = int(self.slab.shape[-1]) * esz
        nrows = self.plane_elems // int(self.slab.shape[-1])
        plane_bytes = self.plane_elems * esz
        early = 0
        if self.lower is not None and self.ndim == 3:
            tiles_x = -(-int(self.slab.shape[-1]) // 32)
            early = min(nrows, (-(-self.lib.sm_count() // tiles_x) + 1) * 64)               # tile rows of the first wave (+1 for the halo rows)
            if early >= nrows:
                early = 0
        if self.lower is not None and self.h_lo > 0:
            if early:
                self.lib.memcpy2d_async(self.halo_lo_ptr, plane_bytes, self.peer_lo_ptr, plane_bytes, early * row_bytes, self.h_lo, side)
                self.lib.memset_async(flags, self._epoch, 1, side)
        if self.upper is not None:
            if self.h_hi > 0:
                self.lib.memcpy_async(self.halo_hi_ptr, self.peer_hi_ptr, self.h_hi * plane_bytes, side)
            self.lib.memset_async(flags + 2, self._epoch, 1, side)
        if self.lower is not None and self.h_lo == 0:
            self.lib.memset_async(flags, self._epoch, 1, side)
            self.lib.memset_async(flags + 1, self._epoch, 1, side)
        if self.lower is not None and self.h_lo > 0:
            if early:
                self.lib.memcpy2d_async(self.halo_lo_ptr + early * row_bytes, plane_bytes, self.peer_lo_ptr + early * row_bytes,
                                        plane_bytes, (nrows - early) * row_bytes, self.h_lo, side)
            else:
                self.lib.memcpy_async(self.halo_lo_ptr, self.peer_lo_ptr, self.h_lo * plane_bytes, side)
            self.lib.memset_async(flags + 1, self._epoch, 1, side)
        self._ev_done.record(self._side)
        try:
            self.lib.imfilter_slab_staged(
                _desc(self.slab, self.n0f8), _desc(self.out), self.stages, self.border.to_abi(self.ndim),
                self.global_planes, self.first,
                self.halo_lo_ptr, self.h_lo if self.lower is not None else 0,
                self.halo_hi_ptr, self.h_hi if self.upper is not None else 0,
                flags, flags + 2, self._epoch, early, stream)
            main.wait_event(self._ev_done)
        except NotSupportedError:
            # not the fused Float32 3-D kernel (or not TMA-capable): same result through direct peer reads from now on
            main.wait_event(self._ev_done)
            self.mode = "p2p"
            self.halo_lo_ptr, self.halo_hi_ptr = self.peer_lo_ptr, self.peer_hi_ptr
            return self.run(sync=False)
        return self.out

    def close(self):
        if self._ctx is not None:
            self.barrier(full=True)             # nobody may still be reading my planes
            self.torch.cuda.synchronize()
            if self._nccl:
                self.dist.barrier(group=self.group)
            self.lib.check(self.lib.dll.b2f_shard_ctx_destroy(self._ctx))
            self._ctx = None
            return
        if self.world > 1 and self.mode in ("p2p", "staged"):
            self.barrier(full=True)             # nobody may still be reading my planes
            if self.slab.is_cuda:
                self.torch.cuda.synchronize()
        for base, off in self._opened:
            self.lib.ipc_close(base, off)
        self._opened = []
        if getattr(self, "_sync_ptr", 0):
            self.lib.free(self._sync_ptr)
            self._sync_ptr = 0


def imfilter_sharded(slab, kernel, border="replicate", *, out=None, group=None, mode="auto", out_dtype=None,
                     n0f8=False, _library=None):
    """One-shot form: plan, run once, release.  Returns this rank's planes of the filtered array."""
    f = ShardedImfilter(slab, kernel, border, out=out, group=group, mode=mode, out_dtype=out_dtype, n0f8=n0f8,
                        _library=_library)
    try:
        res = f.run()
        if slab.is_cuda:
            f.torch.cuda.synchronize()
        return res
    finally:
        f.close()


# ---------------------------------------------------------------------------------------------------------------------
# Slab-sharded mapwindow (one volume too large for one GPU, or already distributed)
# ---------------------------------------------------------------------------------------------------------------------
class ShardedMapwindow:
    """`mapwindow(f, A, window; border)` for f in {extrema, minimum, maximum, median, mean, sum} on an array whose LAST Julia axis
    is partitioned over the ranks (reference src/mapwindow.jl:75-121; the window loop needs, along the sharded axis, the
    `-window_lo` planes below and `window_hi` planes above every owned plane).

    Each rank keeps ONE buffer [lower halo | own planes | upper halo]; `self.slab` is the view of its own planes (fill it in
    place, or pass a tensor to `run`).  `run()` exchanges the raw boundary planes (NCCL send/recv over NVLink into the halo
    parts of the buffer; gloo in the CPU tests), then calls the single-GPU entry point (`b2f_mapwindow_extrema` /
    `b2f_mapwindow_reduce`) on the buffer with the OUTPUT restricted to the owned planes — the kernels evaluate exactly the
    indices of `out`'s axes, so no halo plane is computed twice.  A face of the buffer is either a seam (never reached by an
    owned plane's window) or a face of the whole array (where the border rule then applies unchanged), so the result equals
    the owned planes of `mapwindow` on the whole array.  No collective in the data path.

        f = ShardedMapwindow(ifb.extrema, slab, (7, 7, 7))
        lo, hi = f.run()            # extrema: (min, max); the other window functions: one tensor
    """

    def __init__(self, f, slab, window, border="replicate", *, group=None, _library=None):
        import torch
        import torch.distributed as dist
        from .border import Inner, NoPad
        from .mapwindow import _REDUCE_OP, _kind, _reduce_out_dtype, resolve_window
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if _library is not None:
            self.lib = _library
        else:
            from ._lib import lib
            self.lib = lib()                      # raises when the CUDA extension is missing: no CPU path
        if slab.dim() < 2 or slab.dim() > _abi.MAXDIM:
            raise NotSupportedError("slab-sharded arrays need 2..4 dimensions")
        self.kind = _kind(f)
        self.ndim = slab.dim()
        self.wlo, self.whi = resolve_window(window, self.ndim, allow_even=False)
        b = borderinstance(border)
        if isinstance(b, (Inner, NoPad)) or not isinstance(b, (Pad, Fill)):
            raise NotSupportedError("the sharded mapwindow takes Pad(style) / Fill(value) borders")
        self.border = b
        dts = {torch.uint8: _abi.U8, torch.int16: _abi.I16, torch.int32: _abi.I32, torch.int64: _abi.I64, torch.float32: _abi.F32,
               torch.float64: _abi.F64}
        if slab.dtype not in dts:
            raise NotSupportedError(f"slab eltype {slab.dtype} is not supported")
        self.dt = dts[slab.dtype]
        self.h_lo, self.h_hi = max(0, -self.wlo[-1]), max(0, self.whi[-1])
        n = int(slab.shape[0])
        counts = [n]
        if self.world > 1:
            counts = [None] * self.world
            dist.all_gather_object(counts, n, group=group)
        self.counts = counts
        self.first, self.global_planes = int(sum(counts[:self.rank])), int(sum(counts))
        # no wrap-around neighbour, Pad(:circular) included: the window loop pads the IN-IMAGE PART OF EACH WINDOW
        # (copy_win!, src/mapwindow.jl:310-333), so a window at a face of the array never sees the opposite face
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank + 1 < self.world else None
        if self.h_lo == 0 and self.h_hi == 0:       # symmetric neighbour relation, as in ShardedImfilter
            self.lower = self.upper = None
        if self.world > 1 and min(counts) <= max(self.h_lo, self.h_hi):
            raise DimensionMismatch(f"every rank must own more planes than the window reaches ({max(self.h_lo, self.h_hi)}); counts = {counts}")
        self.n_lo = self.h_lo if self.lower is not None else 0
        self.n_hi = self.h_hi if self.upper is not None else 0
        shp = tuple(slab.shape[1:])
        self.ext = torch.empty((self.n_lo + n + self.n_hi,) + shp, dtype=slab.dtype, device=slab.device)
        self.slab = self.ext[self.n_lo:self.n_lo + n]
        self.slab.copy_(slab)
        self.recv_lo = self.ext[:self.n_lo] if self.n_lo else None
        self.recv_hi = self.ext[self.n_lo + n:] if self.n_hi else None
        self._nccl = self.world > 1 and dist.get_backend(group) == "nccl"
        odt = {v: k for k, v in dts.items()}
        if self.kind in ("extrema", "min", "max"):
            self.out_dt = self.dt
            self._op = None
        else:
            self.out_dt = _reduce_out_dtype(self.kind, self.dt)
            self._op = _REDUCE_OP[self.kind]
        mk = lambda: torch.empty(slab.shape, dtype=odt[self.out_dt], device=slab.device)
        self.out = (mk(), mk()) if self.kind == "extrema" else mk()

    _exchange = ShardedImfilter._exchange
    _grank = ShardedImfilter._grank

    def run(self, slab=None):
        t = self.torch
        if slab is not None:
            self.slab.copy_(slab)
        if self.world > 1:
            self._exchange()
        stream = t.cuda.current_stream().cuda_stream if self.ext.is_cuda else 0
        nd = self.ndim
        mem = _abi.DEVICE if self.ext.is_cuda else _abi.HOST
        img = _abi.make_array(self.ext.data_ptr(), self.dt, tuple(reversed(self.ext.shape)), (1,) * nd, mem)
        odims = tuple(reversed(self.slab.shape))
        ofirst = (1,) * (nd - 1) + (1 + self.n_lo,)                   # the owned planes, in the buffer's coordinates
        od = lambda x: _abi.make_array(x.data_ptr(), self.out_dt, odims, ofirst, mem)
        b = self.border.to_abi(nd)
        if self.kind == "extrema":
            self.lib.mapwindow_extrema(img, od(self.out[0]), od(self.out[1]), False, self.wlo, self.whi, b, stream)
        elif self.kind in ("min", "max"):
            self.lib.mapwindow_extrema(img, od(self.out) if self.kind == "min" else None, od(self.out) if self.kind == "max" else None,
                                       False, self.wlo, self.whi, b, stream)
        else:
            self.lib.mapwindow_reduce(img, od(self.out), self._op, self.wlo, self.whi, b, None, None, stream)
        return self.out


def mapwindow_sharded(f, slab, window, border="replicate", *, group=None, _library=None):
    """One-shot form of ShardedMapwindow: this rank's planes of mapwindow(f, whole, window; border)."""
    m = ShardedMapwindow(f, slab, window, border, group=group, _library=_library)
    res = m.run()
    if slab.is_cuda:
        m.torch.cuda.synchronize()
    return res
