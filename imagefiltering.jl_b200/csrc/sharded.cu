// sharded.cu — the slab-sharded imfilter driver behind the C ABI (b2f_shard_* / b2f_imfilter_sharded, include/b2f.h).
//
// One process (or thread) per GPU holds a slab of planes of the array's LAST axis (BASELINE config 5).  This
// file owns everything a rank does per filter pass besides the compute kernel itself:
//   * exchange buffers: local halo buffers for the `h_lo` / `h_hi` raw boundary planes of the neighbours, flag bytes the
//     kernel polls, and two 32-bit hand-shake words the neighbours write;
//   * the per-pass sequence (all stream-ordered, no host synchronisation, no collective):
//       1. hand-shake: write my pass counter into both neighbours' hand-shake words (a 32-bit write over NVLink) and make my
//          stream wait until theirs have reached mine — "the neighbours' input slabs are complete";
//       2. side stream: the COPY ENGINES pull the neighbours' boundary planes into the local halo buffers (the rows the
//          first wave of tiles reads first, then the upper halo, then the rest), each part followed by its flag byte;
//       3. main stream: the fused kernel starts at once; a CTA waits for a flag only when it is about to read a halo plane;
//       4. the main stream joins the side stream.
// The processes find each other through 128-byte "blobs" (CUDA IPC handles of the slab and of the hand-shake words plus the
// plane count) which the CALLER moves between ranks with whatever it has — MPI, Distributed.jl, a file, torch.distributed:
// the library itself has no communication dependency.  Arrays other than Float32 3-D volumes on the fused kernel, or
// volumes the TMA path cannot take, run the same sequence with direct peer reads (b2f_imfilter_slab) after the hand-shake.
#include <cstring>

#include "common.cuh"

extern "C" {
int b2f_ipc_export(const void *dptr, void *handle64, uint64_t *offset);
int b2f_ipc_open(const void *handle64, uint64_t offset, void **dptr);
int b2f_ipc_close(void *dptr, uint64_t offset);
int b2f_stream_write32(void *dptr, uint32_t value, void *stream);
int b2f_stream_wait_geq32(void *dptr, uint32_t value, void *stream);
}

namespace b2f {

struct ShardBlob {                       // what a rank publishes (B2F_SHARD_BLOB = 256 bytes on the wire)
    uint32_t magic, rank;
    int64_t planes;                      // planes this rank owns
    uint64_t slab_off, sync_off;
    unsigned char slab_handle[64], sync_handle[64];
};
static_assert(sizeof(ShardBlob) <= 256, "blob size");

struct ShardCtx {
    int rank = 0, world = 1;
    const void *slab = nullptr;          // this rank's input slab (fixed at export time: the neighbours map it)
    int64_t planes = 0;
    uint32_t *sync = nullptr;            // [0] written by the lower neighbour, [1] by the upper one
    unsigned char *flags = nullptr;      // lo early, lo rest, hi, pad
    void *recv_lo = nullptr, *recv_hi = nullptr;
    size_t recv_lo_bytes = 0, recv_hi_bytes = 0;
    // neighbours (after connect)
    bool has_lower = false, has_upper = false;
    void *lower_slab = nullptr, *upper_slab = nullptr, *lower_sync = nullptr, *upper_sync = nullptr;
    uint64_t lower_slab_off = 0, upper_slab_off = 0, lower_sync_off = 0, upper_sync_off = 0;
    int64_t lower_planes = 0, upper_planes = 0;
    bool same_peer = false;              // two ranks with a circular wrap: one neighbour on both sides
    cudaStream_t side = nullptr;
    cudaEvent_t ev_go = nullptr, ev_done = nullptr;
    uint32_t step = 0;
    int epoch = 0;
    bool staged_ok = true;
};

}  // namespace b2f

using namespace b2f;

extern "C" {

struct b2f_shard_ctx { ShardCtx c; };

int b2f_shard_ctx_create(b2f_shard_ctx **ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(B2F_EARG, "bad rank / world");
    b2f_shard_ctx *p = new b2f_shard_ctx();
    p->c.rank = rank;
    p->c.world = world;
    cudaError_t e = cudaMalloc((void **)&p->c.sync, 8);
    if (e == cudaSuccess) e = cudaMemset(p->c.sync, 0, 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->c.flags, 4);
    if (e == cudaSuccess) e = cudaMemset(p->c.flags, 0, 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->c.side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->c.ev_go, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->c.ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { delete p; return fail(B2F_ECUDA, "shard context: %s", cudaGetErrorString(e)); }
    *ctx = p;
    return 0;
}

int b2f_shard_ctx_export(b2f_shard_ctx *ctx, const void *slab, int64_t planes, void *blob256) {
    if (!ctx || !slab || !blob256 || planes < 1) return fail(B2F_EARG, "NULL argument");
    ShardBlob b;
    memset(&b, 0, sizeof b);
    b.magic = 0xb2f5b10bu;
    b.rank = (uint32_t)ctx->c.rank;
    b.planes = planes;
    int rc = b2f_ipc_export(slab, b.slab_handle, &b.slab_off);
    if (rc) return rc;
    rc = b2f_ipc_export(ctx->c.sync, b.sync_handle, &b.sync_off);
    if (rc) return rc;
    ctx->c.slab = slab;
    ctx->c.planes = planes;
    memset(blob256, 0, 256);
    memcpy(blob256, &b, sizeof b);
    return 0;
}

static int shard_open(const ShardBlob &b, void **slab, uint64_t *slab_off, void **sync, uint64_t *sync_off) {
    if (b.magic != 0xb2f5b10bu) return fail(B2F_EARG, "not a shard blob");
    int rc = b2f_ipc_open(b.slab_handle, b.slab_off, slab);
    if (rc) return rc;
    *slab_off = b.slab_off;
    rc = b2f_ipc_open(b.sync_handle, b.sync_off, sync);
    if (rc) return rc;
    *sync_off = b.sync_off;
    return 0;
}

int b2f_shard_ctx_connect(b2f_shard_ctx *ctx, const void *lower_blob256, const void *upper_blob256) {
    if (!ctx) return fail(B2F_EARG, "NULL argument");
    ShardCtx &c = ctx->c;
    if (c.has_lower || c.has_upper) return fail(B2F_EARG, "shard context is already connected");
    ShardBlob lo, up;
    if (lower_blob256) memcpy(&lo, lower_blob256, sizeof lo);
    if (upper_blob256) memcpy(&up, upper_blob256, sizeof up);
    if (lower_blob256) {
        int rc = shard_open(lo, &c.lower_slab, &c.lower_slab_off, &c.lower_sync, &c.lower_sync_off);
        if (rc) return rc;
        c.has_lower = true;
        c.lower_planes = lo.planes;
    }
    if (upper_blob256) {
        if (lower_blob256 && lo.rank == up.rank) {      // the same process on both sides (2 ranks, circular): map it once
            c.upper_slab = c.lower_slab; c.upper_sync = c.lower_sync;
            c.upper_slab_off = c.lower_slab_off; c.upper_sync_off = c.lower_sync_off;
            c.same_peer = true;
        } else {
            int rc = shard_open(up, &c.upper_slab, &c.upper_slab_off, &c.upper_sync, &c.upper_sync_off);
            if (rc) return rc;
        }
        c.has_upper = true;
        c.upper_planes = up.planes;
    }
    return 0;
}

int b2f_shard_ctx_destroy(b2f_shard_ctx *ctx) {
    if (!ctx) return 0;
    ShardCtx &c = ctx->c;
    cudaDeviceSynchronize();
    if (c.has_lower) { b2f_ipc_close(c.lower_slab, c.lower_slab_off); b2f_ipc_close(c.lower_sync, c.lower_sync_off); }
    if (c.has_upper && !c.same_peer) { b2f_ipc_close(c.upper_slab, c.upper_slab_off); b2f_ipc_close(c.upper_sync, c.upper_sync_off); }
    if (c.recv_lo) cudaFree(c.recv_lo);
    if (c.recv_hi) cudaFree(c.recv_hi);
    if (c.sync) cudaFree(c.sync);
    if (c.flags) cudaFree(c.flags);
    if (c.side) cudaStreamDestroy(c.side);
    if (c.ev_go) cudaEventDestroy(c.ev_go);
    if (c.ev_done) cudaEventDestroy(c.ev_done);
    delete ctx;
    return 0;
}

// stream-ordered barrier with the neighbours only: "their slabs are complete" / "they are done reading mine"
int b2f_shard_handshake(b2f_shard_ctx *ctx, void *stream) {
    if (!ctx) return fail(B2F_EARG, "NULL argument");
    ShardCtx &c = ctx->c;
    c.step += 1;
    int rc = 0;
    if (c.has_lower && !rc) rc = b2f_stream_write32((char *)c.lower_sync + 4, c.step, stream);     // I am its upper neighbour
    if (c.has_upper && !rc) rc = b2f_stream_write32((char *)c.upper_sync, c.step, stream);         // I am its lower neighbour
    if (c.has_lower && !rc) rc = b2f_stream_wait_geq32(c.sync, c.step, stream);
    if (c.has_upper && !rc) rc = b2f_stream_wait_geq32(c.sync + 1, c.step, stream);
    return rc;
}

int b2f_imfilter_sharded(b2f_shard_ctx *ctx, const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                         const b2f_border *border, int64_t global_last_dim, int64_t slab_first, void *stream) {
    if (!ctx || !img || !out || !border) return fail(B2F_EARG, "NULL argument");
    ShardCtx &c = ctx->c;
    cudaStream_t st = (cudaStream_t)stream;
    if (img->mem != B2F_DEVICE || out->mem != B2F_DEVICE) return fail(B2F_EARG, "the sharded path takes device arrays");
    if (img->ptr != c.slab) return fail(B2F_EARG, "img is not the slab this context exported");
    const int nd = img->ndim;
    if (nd < 2 || nd > B2F_MAXDIM || img->dims[nd - 1] != c.planes) return fail(B2F_EDIM, "slab shape differs from the exported one");
    // halo depth = accumulated reach of the cascade along the last axis (Pad{0}(kernel): src/border.jl:602-642)
    int64_t zfirst = 0, zlast = 0;
    for (int s = 0; s < nstages; ++s) {
        const b2f_stage &k = stages[s];
        if (k.kind == B2F_STAGE_1D && k.axis != nd - 1) continue;
        if (k.kind == B2F_STAGE_LAPLACIAN) { if (k.len[nd - 1] == 3) { zfirst -= 1; zlast += 1; } continue; }
        zfirst += k.lo[nd - 1];
        zlast += k.lo[nd - 1] + k.len[nd - 1] - 1;
    }
    const int64_t h_lo = zfirst < 0 ? -zfirst : 0, h_hi = zlast > 0 ? zlast : 0;
    const bool use_lo = c.has_lower && h_lo > 0, use_hi = c.has_upper && h_hi > 0;
    if (use_lo && c.lower_planes < h_lo) return fail(B2F_EDIM, "the lower neighbour owns fewer planes than the halo needs");
    if (use_hi && c.upper_planes < h_hi) return fail(B2F_EDIM, "the upper neighbour owns fewer planes than the halo needs");
    int64_t plane_elems = 1;
    for (int d = 0; d < nd - 1; ++d) plane_elems *= img->dims[d];
    const size_t esz = dtype_size(img->dtype), plane_bytes = (size_t)plane_elems * esz;
    const char *peer_lo = use_lo ? (const char *)c.lower_slab + (size_t)(c.lower_planes - h_lo) * plane_bytes : nullptr;
    const char *peer_hi = use_hi ? (const char *)c.upper_slab : nullptr;

    int rc = b2f_shard_handshake(ctx, stream);
    if (rc) return rc;
    if (!use_lo && !use_hi)
        return b2f_imfilter_slab(img, out, stages, nstages, border, global_last_dim, slab_first, nullptr, 0, nullptr, 0, stream);

    if (c.staged_ok && nd == 3 && img->dtype == B2F_F32 && out->dtype == B2F_F32) {
        // local halo buffers (grown on demand)
        if (use_lo && c.recv_lo_bytes < (size_t)h_lo * plane_bytes) {
            if (c.recv_lo) cudaFree(c.recv_lo);
            B2F_CUDA(cudaMalloc(&c.recv_lo, (size_t)h_lo * plane_bytes));
            c.recv_lo_bytes = (size_t)h_lo * plane_bytes;
        }
        if (use_hi && c.recv_hi_bytes < (size_t)h_hi * plane_bytes) {
            if (c.recv_hi) cudaFree(c.recv_hi);
            B2F_CUDA(cudaMalloc(&c.recv_hi, (size_t)h_hi * plane_bytes));
            c.recv_hi_bytes = (size_t)h_hi * plane_bytes;
        }
        c.epoch = c.epoch % 255 + 1;
        B2F_CUDA(cudaEventRecord(c.ev_go, st));
        B2F_CUDA(cudaStreamWaitEvent(c.side, c.ev_go, 0));
        const size_t row_bytes = (size_t)img->dims[0] * esz;
        const int64_t nrows = plane_elems / img->dims[0];
        // copy order = order of need: the rows of the lower halo the first wave of tiles reads, the upper halo (read when the
        // first marches end), the rest of the lower halo
        int64_t early = 0;
        if (use_lo) {
            const int64_t tiles_x = (img->dims[0] + 31) / 32;
            early = ((sm_count() + tiles_x - 1) / tiles_x + 1) * 64;
            if (early >= nrows) early = 0;
        }
        if (use_lo && early) {
            B2F_CUDA(cudaMemcpy2DAsync(c.recv_lo, plane_bytes, peer_lo, plane_bytes, (size_t)early * row_bytes, (size_t)h_lo, cudaMemcpyDefault, c.side));
            B2F_CUDA(cudaMemsetAsync(c.flags, c.epoch, 1, c.side));
        }
        if (use_hi) {
            B2F_CUDA(cudaMemcpyAsync(c.recv_hi, peer_hi, (size_t)h_hi * plane_bytes, cudaMemcpyDefault, c.side));
            B2F_CUDA(cudaMemsetAsync(c.flags + 2, c.epoch, 1, c.side));
        }
        if (use_lo) {
            if (early)
                B2F_CUDA(cudaMemcpy2DAsync((char *)c.recv_lo + (size_t)early * row_bytes, plane_bytes, peer_lo + (size_t)early * row_bytes, plane_bytes,
                                           (size_t)(nrows - early) * row_bytes, (size_t)h_lo, cudaMemcpyDefault, c.side));
            else
                B2F_CUDA(cudaMemcpyAsync(c.recv_lo, peer_lo, (size_t)h_lo * plane_bytes, cudaMemcpyDefault, c.side));
            B2F_CUDA(cudaMemsetAsync(c.flags + 1, c.epoch, 1, c.side));
        }
        B2F_CUDA(cudaEventRecord(c.ev_done, c.side));
        rc = b2f_imfilter_slab_staged(img, out, stages, nstages, border, global_last_dim, slab_first, use_lo ? c.recv_lo : nullptr,
                                      use_lo ? h_lo : 0, use_hi ? c.recv_hi : nullptr, use_hi ? h_hi : 0, c.flags, c.flags + 2, c.epoch,
                                      (int32_t)early, stream);
        B2F_CUDA(cudaStreamWaitEvent(st, c.ev_done, 0));
        if (rc != B2F_ENOTSUP) return rc;
        c.staged_ok = false;               // not the fused kernel / not TMA-capable: direct peer reads from now on
    }
    return b2f_imfilter_slab(img, out, stages, nstages, border, global_last_dim, slab_first, peer_lo, use_lo ? h_lo : 0, peer_hi,
                             use_hi ? h_hi : 0, stream);
}

}  // extern "C"
