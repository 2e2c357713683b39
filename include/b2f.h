/* b2f.h — C ABI of the B200-native FIR / running-extrema hot path.
 *
 * This is the drop-in boundary for JuliaImages/ImageFiltering.jl's FIR path.  A Julia shim
 *  reaches these entry points by `ccall` from new methods of
 *     imfilter!(r::CUDALibs{<:Algorithm.FIR}, out, img, kernel::ProcessedKernel, border)
 * which sit in front of the reference's own border/scheduler/loop methods
 * (reference src/imfilter.jl:321-341 [pad], :367-457 [scheduler], :592-739 [loops]) and of
 *     mapwindow(extrema|minimum|maximum, img, window)      (reference src/mapwindow.jl:75-121,337-481).
 *
 * Two libraries export this same ABI:
 *   libb2f.so         — the product: CUDA sm_100a kernels, no CPU execution path at all.
 *   libb2f_oracle.so  — TEST INFRASTRUCTURE: a single-threaded CPU restatement of the reference
 *                       algorithm (oracle/), used only as the parity checker and CPU baseline.
 *
 * Conventions (all follow the reference):
 *   - arrays are dense column-major ("Julia order"): dims[0] is the fastest axis;
 *   - every array carries `origin[d]` = the index of its first element along axis d
 *     (Julia `first(axes(A,d))`; 1 for a plain Array, anything for an OffsetArray);
 *   - filtering is CORRELATION: out[I] = sum_J img[I+J] * kernel[J]   (docs/src/kernels.md:39-50);
 *   - a kernel stage carries `lo[d]` = index of its first tap along axis d (−h for a centred kernel);
 *   - stages of a cascade are applied in array order (src/imfilter.jl:438-446);
 *   - `out`'s axes select the computed region (src/imfilter.jl:604-615);
 *   - semantics of a cascade = "pad the input ONCE by the accumulated extent of the whole
 *     cascade, then run every stage as a valid filter over a shrinking region"
 *     (src/imfilter.jl:331-341,385-395,438-446; src/border.jl:614-642,657-684).
 */
#ifndef B2F_H
#define B2F_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_MAXDIM 4
#define B2F_MAXSTAGES 16

/* element types.  N0F8 = FixedPointNumbers.N0f8: raw byte i means the value i/255
 * (conversion happens when the reference pads: src/border.jl:343). */
enum {
    B2F_U8 = 0,
    B2F_N0F8 = 1,
    B2F_I16 = 2,
    B2F_I32 = 3,
    B2F_I64 = 4,
    B2F_F32 = 5,
    B2F_F64 = 6,
    B2F_U16 = 7,
    B2F_U32 = 8
};

/* where `ptr` lives */
enum { B2F_HOST = 0, B2F_DEVICE = 1 };

/* border styles.  src/border.jl:564-590 (Pad styles), :383-384 (Fill), :547-550 (Inner),
 * src/imfilter.jl:256-278 (NoPad). */
enum {
    B2F_REPLICATE = 0,
    B2F_CIRCULAR = 1,
    B2F_SYMMETRIC = 2,
    B2F_REFLECT = 3,
    B2F_FILL = 4,
    B2F_INNER = 5,
    B2F_NOPAD = 6
};

/* stage kinds: a 1-D factor acting along one axis (KernelFactors.ReshapedOneD,
 * src/kernelfactors.jl:52-131), a dense N-d block (src/imfilter.jl:624-669), or the opaque
 * Kernel.Laplacian stencil (src/kernel.jl:341-391, loop src/specialty.jl:3-16): len[d] = 3 on the
 * flagged axes and 1 elsewhere, lo[d] = -1 / 0, taps ignored. */
enum { B2F_STAGE_1D = 0, B2F_STAGE_DENSE = 1, B2F_STAGE_LAPLACIAN = 2 };

/* tap element type as seen by the reference's accumulator typing (src/imfilter.jl:630-632,
 * src/utils.jl:122-133).  Taps are always PASSED as double. */
enum { B2F_TAPS_F64 = 0, B2F_TAPS_F32 = 1, B2F_TAPS_INT = 2 };

/* status codes; the Julia shim maps them back to the reference's exception types */
enum {
    B2F_OK = 0,
    B2F_EDIM = -1,     /* DimensionMismatch   (src/imfilter.jl:604-615) */
    B2F_EARG = -2,     /* ArgumentError       (src/border.jl:147-155,250-262) */
    B2F_EINEXACT = -3, /* InexactError        (narrowing store, src/imfilter.jl:233-254) */
    B2F_ECUDA = -4,    /* CUDA runtime failure */
    B2F_ENOTSUP = -5,  /* valid reference call that this library does not accelerate */
    B2F_ENOMEM = -6
};

typedef struct {
    void *ptr;
    int32_t dtype;
    int32_t ndim;
    int64_t dims[B2F_MAXDIM];
    int64_t origin[B2F_MAXDIM];
    int32_t mem;
    int32_t reserved;
} b2f_array;

typedef struct {
    int32_t kind;               /* B2F_STAGE_1D | B2F_STAGE_DENSE */
    int32_t axis;               /* 1-D: the axis (0-based) the taps act on */
    int32_t ndim;               /* dense: number of axes of the block (== array ndim) */
    int32_t tap_dtype;          /* B2F_TAPS_* */
    int64_t len[B2F_MAXDIM];    /* taps per axis (1-D: only len[axis] is read) */
    int64_t lo[B2F_MAXDIM];     /* index of the first tap per axis (1-D: only lo[axis]) */
    const double *taps;         /* HOST pointer, column-major, prod(len) values */
} b2f_stage;

typedef struct {
    int32_t style;              /* B2F_REPLICATE … B2F_NOPAD */
    int32_t npad;               /* 0: derive the padding from the kernel (Pad{0}, Fill{T,0}, Inner{0});
                                   ndim: explicit lo/hi below (Pad{N}/Fill{T,N}/Inner{N}) */
    double fill;                /* B2F_FILL: the value, in image-value units */
    int64_t lo[B2F_MAXDIM];
    int64_t hi[B2F_MAXDIM];
} b2f_border;

/* ---- library-level ---------------------------------------------------------------------- */
const char *b2f_version(void);
/* thread-local message for the last non-OK status returned on this thread */
const char *b2f_last_error(void);
/* 1 for libb2f.so (CUDA), 0 for the oracle library */
int b2f_is_device_library(void);

/* ---- device memory helpers (so the Julia side needs no CUDA.jl) --------------------------- */
int b2f_set_device(int device);
int b2f_device_count(int *count);
/* multiprocessors of the current device (grid sizing of the sharded driver; 148 on a B200) */
int b2f_sm_count(int *count);
int b2f_malloc(void **dptr, uint64_t bytes);
int b2f_free(void *dptr);
int b2f_host_alloc(void **hptr, uint64_t bytes);   /* pinned */
int b2f_host_free(void *hptr);
/* Pin an EXISTING host allocation in place (cudaHostRegister) / undo it: a Julia `Array` registered once is copied at
 * full PCIe speed by every later call that passes it as a B2F_HOST array (an unregistered, pageable array is copied by
 * the CUDA runtime through its own bounce buffers: about half the bandwidth, and the copy blocks the calling thread). */
int b2f_host_register(void *hptr, uint64_t bytes);
int b2f_host_unregister(void *hptr);
int b2f_memcpy_h2d(void *dptr, const void *hptr, uint64_t bytes);
int b2f_memcpy_d2h(void *hptr, const void *dptr, uint64_t bytes);
int b2f_sync(void);

/* Peer mapping across processes (one process per GPU): export a 64-byte handle + offset for device memory owned by
 * this process (any pointer inside a cudaMalloc'ed allocation), open it in another process on the same node as a
 * device pointer usable by kernels and copies (NVLink P2P loads), and close it again.  Used by the sharded path so
 * that the filter kernel reads its neighbours' boundary planes directly (b2f_imfilter_slab's halo pointers). */
int b2f_ipc_export(const void *dptr, void *handle64, uint64_t *offset);
int b2f_ipc_open(const void *handle64, uint64_t offset, void **dptr);
int b2f_ipc_close(void *dptr, uint64_t offset);

/* ---- the hot path ------------------------------------------------------------------------ */

/* imfilter!(r, out, img, kernel::ProcessedKernel, border)      replaces src/imfilter.jl:321-341 and
 * everything below it.  `roi_lo/roi_hi` (inclusive index bounds, may be NULL = axes(out)) is the
 * `inds` argument of the NoPad form (src/imfilter.jl:367-395).  `stream` is a cudaStream_t (NULL =
 * default stream).  Host arrays are copied to stream-ordered device allocations and back (cudaMemcpyAsync straight from /
 * to the caller's memory: at full PCIe speed when that memory is pinned — b2f_host_alloc or b2f_host_register — and through
 * the runtime's bounce buffers when it is pageable) and the call is synchronous; device arrays are used in place and the
 * call is asynchronous on `stream`.  A host-to-host call of 64 MiB or more on PINNED arrays whose cascade has a slab form
 * (b2f_imfilter_slab) runs as a three-stream pipeline over chunks of planes of the last axis — upload of chunk c+1, kernel of
 * chunk c, download of chunk c-1 — so both directions of the PCIe link work at once (B2F_HOST_PIPELINE=0 turns it off).
 *
 * Arithmetic follows the reference's typing keyed on eltype(out):
 *   out F64           — double accumulate, separate multiply and add in tap order (bit-exact
 *                       against the oracle whenever every axis is filtered by at most one stage);
 *   out F32           — float FMA accumulate (within 1e-5 * prod_stage(sum|k|) * max|img|);
 *   out integer types — int64 accumulate; needs integer taps and integer input; a result that
 *                       does not fit eltype(out) gives B2F_EINEXACT.                            */
int b2f_imfilter(const b2f_array *img, const b2f_array *out,
                 const b2f_stage *stages, int32_t nstages,
                 const b2f_border *border,
                 const int64_t *roi_lo, const int64_t *roi_hi,
                 void *stream);

/* Accumulate mode of the Float64 arithmetic on the CALLING THREAD (like the current device; default B2F_ACCUM_EXACT):
 *   B2F_ACCUM_EXACT  separate multiply and add, as the reference's `tmp += A[i+j]*k[j]` compiles on the CPU
 *                    (src/imfilter.jl:732-737): Float64 outputs are bit-equal to the reference;
 *   B2F_ACCUM_FMA    the library MAY fuse them (one rounding instead of two, half the FP64 instructions: the reference-typed
 *                    configs — Float64 outputs of non-dyadic taps — are FP64-pipe-bound).  A permission, not an obligation:
 *                    the fused form exists for the streamed 2-D kernel with 3x3 (1 or 2 planes), 5x5, 7x7, 9x9, 13x13 and
 *                    17x17 taps; everything else computes as under B2F_ACCUM_EXACT.  Results differ from the reference's by
 *                    at most the rounding of one product per tap (<= 1e-15 * prod_stage(sum|k|) * max|img|).
 * Returns the previous mode, or a negative status for an unknown mode.  Float32 and integer arithmetic are unaffected. */
#define B2F_ACCUM_EXACT 0
#define B2F_ACCUM_FMA 1
int b2f_set_accum_mode(int32_t mode);

/* imgradients(img, kernelfun, border)  (src/specialty.jl:39-53): `nplanes` independent cascades
 * of `nstages_each` stages, all reading the same `img`; plane p uses
 * stages[p*nstages_each … (p+1)*nstages_each-1] and writes outs[p].  The product library reads
 * img from HBM once for all planes. */
int b2f_imgradients(const b2f_array *img, const b2f_array *outs, int32_t nplanes,
                    const b2f_stage *stages, int32_t nstages_each,
                    const b2f_border *border, void *stream);

/* mapwindow(extrema|minimum|maximum, img, window; border)  (src/mapwindow.jl:75-121,337-481).
 * The window along axis d covers indices [i+win_lo[d], i+win_hi[d]].  out_min / out_max may each
 * be NULL; `interleaved` != 0 writes Tuple{T,T} (min,max) pairs into out_min (out_max ignored),
 * which is the memory layout of the reference's `Array{Tuple{T,T}}` result.
 * Border: any Pad style == truncation of the window at the array ends (src/mapwindow.jl:310-317);
 * B2F_FILL includes the fill value where the window leaves the array; B2F_INNER restricts the
 * outputs to out's axes, which must lie in the interior. */
int b2f_mapwindow_extrema(const b2f_array *img, const b2f_array *out_min, const b2f_array *out_max,
                          int32_t interleaved,
                          const int64_t *win_lo, const int64_t *win_hi,
                          const b2f_border *border, void *stream);

/* mapwindow(median!, img, window; border)  (generic window loop src/mapwindow.jl:270-333 with
 * f = Statistics.median!): NaN if the window holds a NaN, else the middle of the sorted window (x/2 + y/2 of the two
 * middle elements for even lengths).  Output eltype: Float32 for Float32 images, Float64 otherwise.  Window [win_lo,
 * win_hi] per axis, at most 128 elements.  Borders as in copy_win! (:310-333): Pad styles pad the window's in-image
 * part by padindex (the remap is relative to that part, not to the whole image), Fill inserts the value, Inner
 * restricts the outputs to out's axes. */
int b2f_mapwindow_median(const b2f_array *img, const b2f_array *out, const int64_t *win_lo, const int64_t *win_hi,
                         const b2f_border *border, void *stream);

/* mapwindow(f, img, window; border, indices) for the window reductions the reference's tests and benchmarks use
 * (generic window loop src/mapwindow.jl:270-306; goldens test/mapwindow.jl:105-152; workloads benchmark/benchmarks.jl:21-35):
 *   B2F_WIN_MEDIAN  as b2f_mapwindow_median;
 *   B2F_WIN_MEAN    sum of the window in column-major order / length: Float32 for Float32 images, Float64 otherwise;
 *   B2F_WIN_SUM     the same sum: eltype(img) for Float32 / Float64, Int64 for integer images (Julia widens small
 *                   integers to Int / UInt);
 *   B2F_WIN_MIN / B2F_WIN_MAX   eltype(img) (the O(prod(w)) loop; b2f_mapwindow_extrema is the fast path).
 * `idx_first[d]` / `idx_step[d]` are the `indices=` ranges (src/mapwindow.jl:123-131,156-183): output element j along axis
 * d is the window at image index idx_first[d] + j * idx_step[d] (image-axis coordinates, i.e. counted like origin[d]);
 * out->dims[d] positions are evaluated.  NULL: the positions are out's own axes (step 1).  Windows are gathered with
 * copy_win!'s border semantics (:310-333).  Float sums follow the window's memory order; Julia's `sum` may reassociate
 * (SIMD), so bit-level agreement with Julia is unpinned for Float windows — exact for integers. */
enum { B2F_WIN_MEDIAN = 0, B2F_WIN_MEAN = 1, B2F_WIN_SUM = 2, B2F_WIN_MIN = 3, B2F_WIN_MAX = 4 };
int b2f_mapwindow_reduce(const b2f_array *img, const b2f_array *out, int32_t op, const int64_t *win_lo, const int64_t *win_hi,
                         const b2f_border *border, const int64_t *idx_first, const int64_t *idx_step, void *stream);

/* Slab form of a separable cascade, used by the sharded N-d path: the array's LAST axis is
 * partitioned across GPUs.  `img` and `out` hold this rank's owned planes [slab_first, slab_first + dims[ndim-1])
 * of an array whose last axis has `global_last_dim` planes.  `halo_lo` / `halo_hi` point to `n_halo_lo` /
 * `n_halo_hi` RAW input planes (eltype(img), dense, same plane shape) lying logically just below / above the owned
 * planes: receive buffers filled by an NCCL send/recv, or the neighbouring GPU's memory mapped into this process
 * (cudaIpcOpenMemHandle) — in which case the kernel performs the halo exchange itself with P2P loads over NVLink.
 * At a global face pass 0 planes: the border style is applied there in GLOBAL plane coordinates (a Pad(:circular)
 * wrap is passed as an ordinary halo).  Result = the owned planes of b2f_imfilter on the whole array
 * (reference semantics: src/imfilter.jl:321-341 pad once, :385-395 cascade).  Device arrays only; asynchronous. */
int b2f_imfilter_slab(const b2f_array *img, const b2f_array *out,
                      const b2f_stage *stages, int32_t nstages,
                      const b2f_border *border,
                      int64_t global_last_dim, int64_t slab_first,
                      const void *halo_lo, int64_t n_halo_lo,
                      const void *halo_hi, int64_t n_halo_hi,
                      void *stream);

/* Staged form of b2f_imfilter_slab: halo_lo / halo_hi are LOCAL buffers that are still being filled when the call is
 * made — typically by b2f_memcpy_async from the neighbour's mapped memory on another stream, each copy followed by
 * b2f_memset_async(flag, epoch, 1, that_stream).  The kernel starts immediately; a CTA waits for `*flag_lo == epoch`
 * (`*flag_hi`) only when it is about to read the first plane of that halo, so the NVLink transfer proceeds at copy
 * speed underneath the computation.  The lower halo is split: flag_lo[0] covers rows [0, lo_early_rows) of every
 * plane (what the first wave of tiles reads; copy it first), flag_lo[1] the remaining rows; lo_early_rows = 0 puts
 * everything under flag_lo[1].  epoch in 1..255, different from the value the flags hold from the previous call.
 * Fused Float32 3-D path only (B2F_ENOTSUP otherwise: use b2f_imfilter_slab). */
int b2f_imfilter_slab_staged(const b2f_array *img, const b2f_array *out,
                             const b2f_stage *stages, int32_t nstages,
                             const b2f_border *border,
                             int64_t global_last_dim, int64_t slab_first,
                             const void *halo_lo, int64_t n_halo_lo,
                             const void *halo_hi, int64_t n_halo_hi,
                             const void *flag_lo, const void *flag_hi, int32_t epoch, int32_t lo_early_rows,
                             void *stream);
/* xy-filtered form of the staged slab call (what b2f_imfilter_sharded uses on the fused Float32 3-D path): instead of RAW halo
 * planes the ranks exchange boundary planes that have already been filtered along every axis but the last — a rank runs
 * b2f_imfilter with the stages of the other axes over its own first / last boundary planes before the pass, the neighbours copy
 * them, and the march of the fused kernel reads both its own boundary planes and the neighbours' in its last-axis stage only.
 * No rank filters a halo plane a second time (with raw halos a slab of n planes pays for n + h_lo + h_hi planes of the other
 * stages) and the result is bit-identical to the unsharded call.
 *   xy_lo: lo_halo + lo_own planes = the planes [slab_first - lo_halo, slab_first + lo_own) of the array, filtered;
 *   xy_hi: hi_own + hi_halo planes = the planes [slab_first + own - hi_own, slab_first + own + hi_halo).
 * lo_halo / hi_halo = 0 at a global face (the border style is applied there in global plane coordinates).  The halo parts may
 * still be on their way: flag_lo[0] / flag_lo[1] / flag_hi, epoch and lo_early_rows as in b2f_imfilter_slab_staged (NULL flags:
 * the planes are there).  Fused Float32 3-D path with TMA only (B2F_ENOTSUP otherwise: use the raw-halo forms). */
typedef struct b2f_slab_xy {
    const void *xy_lo;
    int64_t lo_halo, lo_own;
    const void *xy_hi;
    int64_t hi_own, hi_halo;
} b2f_slab_xy;
int b2f_imfilter_slab_xy(const b2f_array *img, const b2f_array *out,
                         const b2f_stage *stages, int32_t nstages,
                         const b2f_border *border,
                         int64_t global_last_dim, int64_t slab_first,
                         const b2f_slab_xy *xy,
                         const void *flag_lo, const void *flag_hi, int32_t epoch, int32_t lo_early_rows,
                         void *stream);

/* ---- the sharded driver: one rank per GPU, slabs along the last axis ---------------------------------------
 * b2f_imfilter_sharded is b2f_imfilter for ONE RANK's slab of an array partitioned over `world` GPUs of a node: the
 * result equals this rank's planes of the filter applied to the whole array.  Per pass the library performs the neighbour
 * hand-shake (32-bit stream-ordered writes / waits over NVLink), pulls the neighbours' boundary planes with the copy engines
 * into local halo buffers while the fused kernel is already marching, and launches the kernel — no host synchronisation,
 * no collective, no dependency on a communication library.  Set-up:
 *     b2f_shard_ctx_create(&ctx, rank, world);
 *     b2f_shard_ctx_export(ctx, slab_ptr, planes, blob);        // 256-byte blob: CUDA IPC handles + plane count
 *     ... the caller moves the blobs between ranks (MPI_Sendrecv, Distributed.jl, a file, torch.distributed) ...
 *     b2f_shard_ctx_connect(ctx, lower_blob_or_NULL, upper_blob_or_NULL);   (NULL at a global face; under
 *                                                                           Pad(:circular) the wrap-around ranks are neighbours too)
 *     b2f_imfilter_sharded(ctx, img (its ptr == slab_ptr), out, stages, nstages, border, global_last_dim, slab_first, stream);
 * Every rank must make the same sequence of b2f_imfilter_sharded / b2f_shard_handshake calls.  Before REFILLING a slab call
 * b2f_shard_handshake(ctx, stream) (the neighbours have finished reading it once it completes).  Device arrays only. */
#define B2F_SHARD_BLOB 256
typedef struct b2f_shard_ctx b2f_shard_ctx;
int b2f_shard_ctx_create(b2f_shard_ctx **ctx, int32_t rank, int32_t world);
int b2f_shard_ctx_export(b2f_shard_ctx *ctx, const void *slab, int64_t planes, void *blob256);
int b2f_shard_ctx_connect(b2f_shard_ctx *ctx, const void *lower_blob256, const void *upper_blob256);
int b2f_shard_handshake(b2f_shard_ctx *ctx, void *stream);
int b2f_imfilter_sharded(b2f_shard_ctx *ctx, const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                         const b2f_border *border, int64_t global_last_dim, int64_t slab_first, void *stream);
int b2f_shard_ctx_destroy(b2f_shard_ctx *ctx);

/* stream-ordered device copies / byte fills (peer-mapped pointers allowed): the halo staging of the sharded path */
int b2f_memcpy_async(void *dst, const void *src, uint64_t bytes, void *stream);
int b2f_memcpy2d_async(void *dst, uint64_t dpitch, const void *src, uint64_t spitch, uint64_t width, uint64_t height,
                       void *stream);
int b2f_memset_async(void *dptr, int32_t byte, uint64_t bytes, void *stream);
/* stream-ordered 32-bit signal / wait on device memory (peer-mapped pointers allowed): the neighbour hand-shake of the
 * sharded path (a rank writes its step counter into its neighbours' flag words and waits until its own are >= it) */
int b2f_stream_write32(void *dptr, uint32_t value, void *stream);
int b2f_stream_wait_geq32(void *dptr, uint32_t value, void *stream);

/* ---- consumers of the LoG path ------------------------------------------ */

/* findlocalmaxima(img; window, edges) / findlocalminima  (src/extrema.jl:107-164, loop :125-162).
 * `window[d]` >= 1 per axis (half-width window[d] >> 1), `edges[d]` == 0 excludes the first and last
 * index along axis d.  An element is an extremum when it compares strictly greater (minima != 0:
 * smaller) than every neighbour of the window that lies inside the array.  `idx` (HOST, capacity
 * `cap`, may be NULL when cap == 0) receives the first `cap` peaks as 0-based column-major linear
 * indices in ascending order — the order in which the reference pushes its CartesianIndex values;
 * `*count` receives the total number of peaks (call again with cap >= count if it was larger).
 * Synchronous. */
int b2f_findlocalextrema(const b2f_array *img, int32_t minima, const int64_t *window, const int32_t *edges,
                         int64_t *idx, int64_t cap, int64_t *count, void *stream);

/* multiLoG (src/extrema.jl:94-105): `stack` has one LEADING axis more than `src` (dims (S, dims(src)...),
 * Float32 or Float64); slice `slice` (0-based) of that axis receives src .* scale, the product taken in
 * Float64 and stored as eltype(stack) — the reference's `imfilter!(view(img_LoG, isigma, :, ...), ...)`
 * followed by `LoG_slice .*= -σ`, with src = the dense imfilter result.  Device arrays (the oracle
 * library: host arrays); asynchronous on `stream`. */
int b2f_scale_into_slice(const b2f_array *src, const b2f_array *stack, int64_t slice, double scale, void *stream);

/* maximum(abs, img) (src/extrema.jl:85) -> *result (HOST).  NaN if any element is NaN.  Synchronous. */
int b2f_maxabs(const b2f_array *img, double *result, void *stream);

/* values[i] = arr[idx[i]] as Float64, idx = 0-based linear indices (HOST in, HOST out): the amplitudes
 * img_LoG[x] of the peaks (src/extrema.jl:86-90).  Synchronous. */
int b2f_gather(const b2f_array *arr, const int64_t *idx, int64_t n, double *values, void *stream);

/* ---- NA() border: the element-wise pieces around the two FIR calls -------------
 * imfilter(img, kernel, NA(na)) (src/imfilter.jl:282-318) is: flag the NA elements; if the kernel is separable
 * and nothing is flagged, filter with Fill(0) and divide by the per-axis responses to a vector of ones
 * (imfilter_na_separable!, normalize_separable!/normalize_dims! :1123-1127,1234-1250); otherwise filter the image
 * with the flagged elements zeroed and the validity mask, both with Fill(0), and divide (imfilter_na_inseparable!
 * :1110-1121).  The FIR calls are b2f_imfilter; these three are the rest. */

/* naflag = na.(img): na_mode 0 = isnan, 1 = !isfinite, 2 = never.  imgtmp (may be NULL) receives img with the
 * flagged elements zeroed, converted to its eltype (F32/F64); valid (may be NULL) receives !naflag as 0/1 (F32/F64);
 * *hasna = any(naflag).  Synchronous. */
int b2f_na_prepare(const b2f_array *img, int32_t na_mode, const b2f_array *imgtmp, const b2f_array *valid, int32_t *hasna,
                   void *stream);
/* out[I] /= den[I] (src/imfilter.jl:1117-1119): Float32/Float32 divides in Float32, otherwise in Float64, stored as
 * eltype(out). */
int b2f_divide(const b2f_array *out, const b2f_array *den, void *stream);
/* normalize_dims!(A, factors) (src/imfilter.jl:1241-1250): A[I] = (A[I] / f1[I1]) / f2[I2] ... with one HOST Float64
 * vector of length dims[d] per axis.  Synchronous. */
int b2f_normalize_dims(const b2f_array *out, const double *const *factors, void *stream);

/* ---- IIR filtering -----------------------------------------------------------------------------------
 * imfilter!(r, out, img, kernel::TriggsSdika, dim, border)   replaces src/imfilter.jl:922-1092 (_imfilter_dim!, leftborder!,
 * rightborder!, rightΔu): a Young / van Vliet recursive filter along ONE axis with Triggs-Sdika boundary conditions —
 * forward recursion u[i] = x[i] + a1 u[i-1] + a2 u[i-2] + a3 u[i-3] started from the steady state of the value left of the
 * line, backward recursion v[i] = u[i] + b1 v[i+1] + b2 v[i+2] + b3 v[i+3] started from M (u[n-1..n-3] - u+) + v+, final
 * scaling.  `coef` holds 18 doubles: a[3], b[3], scale, M[9] (row-major), 1 - sum(a), 1 - sum(b) (the last two formed in the
 * kernel's own float type, as the reference does).  `border` is B2F_REPLICATE (the line's first / last element continues) or
 * B2F_FILL (border->fill continues); anything else is B2F_EARG ("only replicate is supported", src/imfilter.jl:897).  out is
 * F32 or F64 and every operation (separate multiplies and adds, in the reference's order) is carried out in eltype(out); img
 * may be any real eltype and may BE out (the filter is in-place safe, :890).  A line of 3 or fewer elements is B2F_EDIM
 * (:981-983).  The cascade over several axes (IIRGaussian((s1, s2, ...))) is one call per axis, the later ones in place on
 * out (_imfilter_inplace_tuple!, :946-960).  A kernel that is a copy (all a, b zero and scale 1, :1254) copies. */
int b2f_iir(const b2f_array *img, const b2f_array *out, int32_t axis, const double *coef, const b2f_border *border, void *stream);

/* ---- FFT filtering -----------------------------------------------------------------------------------
 * imfilter!(r::AbstractResource{FFT}, out, img, kernel, border)   replaces src/imfilter.jl:776-888: the image is padded by the
 * border rule (on the device), the kernel — ONE dense stage, kernelconv(kernel...) of a factored kernel, src/imfilter.jl:1257-1280
 * — is placed in a zero array of the padded size with periodic indexing, and out = irfft(rfft(A) .* conj(rfft(krn))) restricted
 * to the requested indices.  The transforms are cuFFT's (loaded on first use; B2F_ENOTSUP when the library is missing), padding,
 * kernel placement, the spectral product and the crop are this library's kernels.  out is F32 or F64 (= the arithmetic type;
 * an integer out is B2F_EINEXACT, as the reference's copy into an Int array is); borders as b2f_imfilter (Pad styles, Fill,
 * Inner); up to 3 transformed axes, later axes the kernel does not touch are a batch.  The result equals b2f_imfilter's up to
 * the rounding of the transforms (the reference asserts `≈` between its FIR and FFT algorithms, test/2d.jl:69-140). */
int b2f_imfilter_fft(const b2f_array *img, const b2f_array *out, const b2f_stage *kernel, const b2f_border *border,
                     const int64_t *roi_lo, const int64_t *roi_hi, void *stream);

/* Measured FP32 multiply-add peak of the current GPU in TFMA/s (a pure fma.rn.f32x2 loop, best of 3, CUDA-event timed): the
 * denominator bench.py reports the dense-kernel path against.  Synchronous.  (The oracle library returns 0.) */
int b2f_bench_fma_peak(double *tfma_per_s, void *stream);

/* number of CUDA kernels launched by this library on the calling thread since the last reset
 * (the oracle library always reports 0) */
int64_t b2f_launch_count(void);
void b2f_reset_launch_count(void);
/* name of the kernel family the last b2f_imfilter/b2f_imgradients/b2f_mapwindow_extrema call
 * on this thread dispatched to ("sep2d", "sep3d", "dense2d", "generic", "extrema", …) */
const char *b2f_last_path(void);

#ifdef __cplusplus
}
#endif
#endif /* B2F_H */
