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
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

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
    cudaStream_t side = nullptr, side2 = nullptr;       // copy streams: lower halo / upper halo (different peers: they run concurrently)
    cudaEvent_t ev_go = nullptr, ev_done = nullptr, ev_done2 = nullptr;
    uint32_t step = 0;
    int epoch = 0;
    bool staged_ok = true;
    // xy-filtered exchange (fused Float32 3-D path): two parity blocks of [h_lo halo | h_hi own | h_lo own | h_hi halo] planes;
    // the neighbours map the whole allocation (its IPC handle travels through the hand-shake buffer, see xy_setup)
    bool xy_ok = true;
    void *xy = nullptr;
    size_t xy_bytes = 0, xy_plane_bytes = 0;
    int64_t xy_hlo = 0, xy_hhi = 0;
    void *lower_xy = nullptr, *upper_xy = nullptr;
    uint64_t lower_xy_off = 0, upper_xy_off = 0;
    uint32_t xy_calls = 0;
};

struct XyPublish {                       // at byte 64 of the hand-shake buffer
    unsigned char handle[64];
    uint64_t off, bytes;
};
constexpr size_t SYNC_BYTES = 256;

}  // namespace b2f

using namespace b2f;

extern "C" {

struct b2f_shard_ctx { ShardCtx c; };

int b2f_shard_ctx_create(b2f_shard_ctx **ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(B2F_EARG, "bad rank / world");
    b2f_shard_ctx *p = new b2f_shard_ctx();
    p->c.rank = rank;
    p->c.world = world;
    cudaError_t e = cudaMalloc((void **)&p->c.sync, SYNC_BYTES);
    if (e == cudaSuccess) e = cudaMemset(p->c.sync, 0, SYNC_BYTES);
    if (e == cudaSuccess) e = cudaMalloc((void **)&p->c.flags, 4);
    if (e == cudaSuccess) e = cudaMemset(p->c.flags, 0, 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->c.side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->c.ev_go, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->c.ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->c.side2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->c.ev_done2, cudaEventDisableTiming);
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
    if (c.lower_xy) b2f_ipc_close(c.lower_xy, c.lower_xy_off);
    if (c.upper_xy && !c.same_peer) b2f_ipc_close(c.upper_xy, c.upper_xy_off);
    if (c.xy) cudaFree(c.xy);
    if (c.recv_lo) cudaFree(c.recv_lo);
    if (c.recv_hi) cudaFree(c.recv_hi);
    if (c.sync) cudaFree(c.sync);
    if (c.flags) cudaFree(c.flags);
    if (c.side) cudaStreamDestroy(c.side);
    if (c.ev_go) cudaEventDestroy(c.ev_go);
    if (c.ev_done) cudaEventDestroy(c.ev_done);
    if (c.side2) cudaStreamDestroy(c.side2);
    if (c.ev_done2) cudaEventDestroy(c.ev_done2);
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

// (Re)allocate the xy exchange buffer and map the neighbours' — collective over the neighbours: every rank reaches it in the
// same call (same kernel, same plane size).  The IPC handle is published in this rank's hand-shake buffer, which the
// neighbours have mapped since connect; a hand-shake orders "published" before "read".
static int xy_setup(b2f_shard_ctx *ctx, size_t plane_bytes, int64_t h_lo, int64_t h_hi, void *stream) {
    ShardCtx &c = ctx->c;
    const size_t need = 2 * (size_t)(2 * (h_lo + h_hi)) * plane_bytes;
    if (c.xy && c.xy_bytes == need && c.xy_plane_bytes == plane_bytes && c.xy_hlo == h_lo && c.xy_hhi == h_hi) return 0;
    B2F_CUDA(cudaDeviceSynchronize());
    if (c.lower_xy) { b2f_ipc_close(c.lower_xy, c.lower_xy_off); c.lower_xy = nullptr; }
    if (c.upper_xy) { if (!c.same_peer) b2f_ipc_close(c.upper_xy, c.upper_xy_off); c.upper_xy = nullptr; }
    if (c.xy) { cudaFree(c.xy); c.xy = nullptr; }
    B2F_CUDA(cudaMalloc(&c.xy, need));
    B2F_CUDA(cudaMemset(c.xy, 0, need));
    c.xy_bytes = need; c.xy_plane_bytes = plane_bytes; c.xy_hlo = h_lo; c.xy_hhi = h_hi;
    XyPublish pub;
    memset(&pub, 0, sizeof pub);
    int rc = b2f_ipc_export(c.xy, pub.handle, &pub.off);
    if (rc) return rc;
    pub.bytes = need;
    B2F_CUDA(cudaMemcpy((char *)c.sync + 64, &pub, sizeof pub, cudaMemcpyHostToDevice));
    B2F_CUDA(cudaDeviceSynchronize());
    rc = b2f_shard_handshake(ctx, stream);
    if (rc) return rc;
    B2F_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    auto open_peer = [&](void *peer_sync, void **xy, uint64_t *off) -> int {
        XyPublish q;
        B2F_CUDA(cudaMemcpy(&q, (char *)peer_sync + 64, sizeof q, cudaMemcpyDeviceToHost));
        if (q.bytes != need) return fail(B2F_EDIM, "the neighbour's xy exchange buffer has a different size (different kernel or plane size?)");
        *off = q.off;
        return b2f_ipc_open(q.handle, q.off, xy);
    };
    if (c.has_lower && (rc = open_peer(c.lower_sync, &c.lower_xy, &c.lower_xy_off))) return rc;
    if (c.has_upper) {
        if (c.same_peer) { c.upper_xy = c.lower_xy; c.upper_xy_off = c.lower_xy_off; }
        else if ((rc = open_peer(c.upper_sync, &c.upper_xy, &c.upper_xy_off))) return rc;
    }
    return 0;
}

// One pass with xy-filtered boundary planes (include/b2f.h, b2f_imfilter_slab_xy): filter my first h_hi and last h_lo planes
// along x and y, hand-shake, let the copy engines pull the neighbours' while the march is already running.
static int sharded_xy(b2f_shard_ctx *ctx, const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                      const b2f_border *border, int64_t global_last_dim, int64_t slab_first, int64_t h_lo, int64_t h_hi,
                      bool use_lo, bool use_hi, void *stream) {
    ShardCtx &c = ctx->c;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t W = img->dims[0], H = img->dims[1], own_n = img->dims[2];
    const size_t plane_bytes = (size_t)W * H * 4, row_bytes = (size_t)W * 4;
    int rc = xy_setup(ctx, plane_bytes, h_lo, h_hi, stream);
    if (rc) return rc;
    const int64_t block = 2 * (h_lo + h_hi);                  // planes of one parity block: [h_lo | h_hi | h_lo | h_hi]
    const int q = (int)(c.xy_calls++ & 1);
    char *mine = (char *)c.xy + (size_t)q * block * plane_bytes;
    char *lo_halo = mine, *lo_own = mine + (size_t)h_lo * plane_bytes, *hi_own = lo_own + (size_t)h_hi * plane_bytes,
         *hi_halo = hi_own + (size_t)h_lo * plane_bytes;
    // 1. stages of the other axes over my boundary planes (the lower neighbour needs my first h_hi planes, the upper one my last h_lo)
    std::vector<b2f_stage> sxy;
    for (int i = 0; i < nstages; ++i)
        if (!(stages[i].kind == B2F_STAGE_1D && stages[i].axis == 2)) sxy.push_back(stages[i]);
    auto prefilter = [&](int64_t first, int64_t n, char *dst) -> int {
        if (n <= 0) return 0;
        b2f_array a = *img, o = *out;
        a.ptr = (char *)img->ptr + (size_t)first * plane_bytes;
        o.ptr = dst;
        a.dims[2] = o.dims[2] = n;
        return b2f_imfilter(&a, &o, sxy.data(), (int32_t)sxy.size(), border, nullptr, nullptr, stream);
    };
    const int64_t n_lo_own = std::min<int64_t>(h_hi, own_n), n_hi_own = std::min<int64_t>(h_lo, own_n);
    if ((rc = prefilter(0, n_lo_own, lo_own))) return rc;
    if ((rc = prefilter(own_n - n_hi_own, n_hi_own, hi_own))) return rc;
    // 2. the neighbours' boundary planes are filtered (and their previous pass no longer reads my other parity block)
    if ((rc = b2f_shard_handshake(ctx, stream))) return rc;
    // 3. copy engines: the rows of the lower halo the first wave of tiles reads, the upper halo, the rest of the lower halo
    c.epoch = c.epoch % 255 + 1;
    B2F_CUDA(cudaEventRecord(c.ev_go, st));
    B2F_CUDA(cudaStreamWaitEvent(c.side, c.ev_go, 0));
    const char *peer_lo = use_lo ? (const char *)c.lower_xy + (size_t)q * block * plane_bytes + (size_t)(h_lo + h_hi) * plane_bytes : nullptr;
    const char *peer_hi = use_hi ? (const char *)c.upper_xy + (size_t)q * block * plane_bytes + (size_t)h_lo * plane_bytes : nullptr;
    int64_t early = 0;
    if (use_lo) {
        const int64_t tiles_x = (W + 31) / 32;
        early = ((sm_count() + tiles_x - 1) / tiles_x + 1) * 64;
        if (early >= H) early = 0;
    }
    if (use_lo && early) {
        B2F_CUDA(cudaMemcpy2DAsync(lo_halo, plane_bytes, peer_lo, plane_bytes, (size_t)early * row_bytes, (size_t)h_lo, cudaMemcpyDefault, c.side));
        B2F_CUDA(cudaMemsetAsync(c.flags, c.epoch, 1, c.side));
    }
    if (use_hi) {
        B2F_CUDA(cudaMemcpyAsync(hi_halo, peer_hi, (size_t)h_hi * plane_bytes, cudaMemcpyDefault, c.side));
        B2F_CUDA(cudaMemsetAsync(c.flags + 2, c.epoch, 1, c.side));
    }
    if (use_lo) {
        if (early)
            B2F_CUDA(cudaMemcpy2DAsync(lo_halo + (size_t)early * row_bytes, plane_bytes, peer_lo + (size_t)early * row_bytes, plane_bytes,
                                       (size_t)(H - early) * row_bytes, (size_t)h_lo, cudaMemcpyDefault, c.side));
        else
            B2F_CUDA(cudaMemcpyAsync(lo_halo, peer_lo, (size_t)h_lo * plane_bytes, cudaMemcpyDefault, c.side));
        B2F_CUDA(cudaMemsetAsync(c.flags + 1, c.epoch, 1, c.side));
    }
    B2F_CUDA(cudaEventRecord(c.ev_done, c.side));
    // 4. the march
    b2f_slab_xy xy;
    xy.xy_lo = use_lo ? lo_halo : lo_own;  xy.lo_halo = use_lo ? h_lo : 0;  xy.lo_own = n_lo_own;
    xy.xy_hi = hi_own;                     xy.hi_own = n_hi_own;            xy.hi_halo = use_hi ? h_hi : 0;
    rc = b2f_imfilter_slab_xy(img, out, stages, nstages, border, global_last_dim, slab_first, &xy, use_lo ? c.flags : nullptr,
                              use_hi ? c.flags + 2 : nullptr, c.epoch, (int32_t)early, stream);
    B2F_CUDA(cudaStreamWaitEvent(st, c.ev_done, 0));
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

    // fused Float32 3-D kernel, B2F_SHARD_XY=1: exchange xy-filtered boundary planes instead of raw ones.  OFF by default: it
    // removes the x / y stages of the 2 x 8 halo planes from the march (-7 % of its multiply-adds on a 128-plane slab) but puts
    // two pre-filter launches and their hand-shake in front of it, and measured slower on 1024^3 (2 GPUs: 1.579 vs 1.506 ms,
    // 8 GPUs: 0.598 vs 0.499 ms; profiles/r2_sharded_xy_ab.jsonl).  The decision depends on the environment, the kernel and the
    // plane shape only, so every rank of the pass takes the same branch.
    static const bool xy_env = getenv("B2F_SHARD_XY") && atoi(getenv("B2F_SHARD_XY")) == 1;
    // (two-sided cascades only: with a one-sided one the last rank consumes no halo but would still have to publish its
    // boundary planes — h_lo / h_hi are properties of the kernel, so this test, too, is the same on every rank)
    if (xy_env && c.xy_ok && (use_lo || use_hi) && h_lo > 0 && h_hi > 0 && nd == 3 && img->dtype == B2F_F32 && out->dtype == B2F_F32) {
        b2f_array gi = *img, go = *out;
        gi.dims[2] = go.dims[2] = global_last_dim;
        gi.ptr = go.ptr = nullptr;
        Plan P;
        if (make_plan(&gi, &go, stages, nstages, border, nullptr, nullptr, P) == 0 && !P.img_ax.empty() && stream3d_xy_capable(P)) {
            int rc = sharded_xy(ctx, img, out, stages, nstages, border, global_last_dim, slab_first, h_lo, h_hi, use_lo, use_hi, stream);
            if (rc != B2F_ENOTSUP) return rc;
            // this rank's slab cannot take the TMA path (alignment): it has published its planes like everybody else and reads
            // the neighbours' RAW planes itself (their slabs are complete: the hand-shake is behind us)
            return b2f_imfilter_slab(img, out, stages, nstages, border, global_last_dim, slab_first, peer_lo, use_lo ? h_lo : 0, peer_hi,
                                     use_hi ? h_hi : 0, stream);
        }
    }
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
        // the upper halo comes from another GPU than the lower one: its copy runs on a second stream, concurrently (one stream
        // would serialise 2 x 32 MB behind each other: ~100 us at link speed, as long as the first wave's march on a thin slab)
        static const bool two_streams = !(getenv("B2F_SHARD_STREAMS") && atoi(getenv("B2F_SHARD_STREAMS")) == 1);
        cudaStream_t hs = two_streams ? c.side2 : c.side;
        if (two_streams && use_hi) B2F_CUDA(cudaStreamWaitEvent(c.side2, c.ev_go, 0));
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
            B2F_CUDA(cudaMemcpyAsync(c.recv_hi, peer_hi, (size_t)h_hi * plane_bytes, cudaMemcpyDefault, hs));
            B2F_CUDA(cudaMemsetAsync(c.flags + 2, c.epoch, 1, hs));
            if (two_streams) B2F_CUDA(cudaEventRecord(c.ev_done2, c.side2));
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
        if (two_streams && use_hi) B2F_CUDA(cudaStreamWaitEvent(st, c.ev_done2, 0));
        if (rc != B2F_ENOTSUP) return rc;
        c.staged_ok = false;               // not the fused kernel / not TMA-capable: direct peer reads from now on
    }
    return b2f_imfilter_slab(img, out, stages, nstages, border, global_last_dim, slab_first, peer_lo, use_lo ? h_lo : 0, peer_hi,
                             use_hi ? h_hi : 0, stream);
}

}  // extern "C"
