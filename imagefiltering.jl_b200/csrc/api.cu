// api.cu — the C ABI of libb2f.so (include/b2f.h): validation, planning, host staging, dispatch.
//
// Replaces, behind `imfilter!(r::CUDALibs{<:FIR}, out, img, kernel, border)`:
//   border resolution + padarray     reference src/imfilter.jl:321-341, src/border.jl:236-352
//   the NoPad "scheduler"            reference src/imfilter.jl:367-457
//   validation                       reference src/imfilter.jl:592-617
// Nothing here computes filter results on the CPU: every path ends in a CUDA kernel launch or an
// error status.  (The padded copy of the reference is never materialised: borders are index
// remapping inside the kernels.)

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace b2f {

static thread_local std::string g_err;
static thread_local std::string g_path = "none";
static thread_local int64_t g_launches = 0;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void set_path(const char *name) { g_path = name; }
void count_launch(int n) { g_launches += n; }
static thread_local int g_accum = B2F_ACCUM_EXACT;
int accum_mode() { return g_accum; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

static int analyse_stage(const b2f_stage *s, int ndim, StageInfo &si) {
    si.s = s;
    for (int d = 0; d < B2F_MAXDIM; ++d) si.lo[d] = si.hi[d] = 0;
    si.copy = false;
    if (s->kind == B2F_STAGE_LAPLACIAN) {
        if (s->ndim != ndim) return fail(B2F_EDIM, "Laplacian has %d dims, array has %d", s->ndim, ndim);
        for (int d = 0; d < ndim; ++d)
            if (s->len[d] == 3) { si.lo[d] = -1; si.hi[d] = 1; }
        return 0;
    }
    if (!s->taps) return fail(B2F_EARG, "stage has NULL taps");
    if (s->kind == B2F_STAGE_1D) {
        if (s->axis < 0 || s->axis >= ndim)
            return fail(B2F_EDIM, "1-D stage axis %d out of range for %d-d array", s->axis, ndim);
        if (s->len[s->axis] < 1) return fail(B2F_EARG, "empty kernel factor");
        si.lo[s->axis] = s->lo[s->axis];
        si.hi[s->axis] = s->lo[s->axis] + s->len[s->axis] - 1;
    } else if (s->kind == B2F_STAGE_DENSE) {
        if (s->ndim != ndim) return fail(B2F_EDIM, "dense kernel has %d dims, array has %d", s->ndim, ndim);
        for (int d = 0; d < ndim; ++d) {
            if (s->len[d] < 1) return fail(B2F_EARG, "empty kernel");
            si.lo[d] = s->lo[d];
            si.hi[d] = s->lo[d] + s->len[d] - 1;
        }
    } else {
        return fail(B2F_EARG, "unknown stage kind %d", s->kind);
    }
    bool unit = true;
    for (int d = 0; d < B2F_MAXDIM; ++d) unit = unit && si.lo[d] == 0 && si.hi[d] == 0;
    si.copy = unit && s->taps[0] == 1.0;  // iscopy, src/imfilter.jl:1252-1255
    return 0;
}

int make_plan(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int nstages,
              const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi, Plan &P) {
    if (!img || !out || !border) return fail(B2F_EARG, "NULL argument");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM)
        return fail(B2F_ENOTSUP, "ndim %d not supported (1..%d)", img->ndim, B2F_MAXDIM);
    if (out->ndim != img->ndim) return fail(B2F_EDIM, "out has %d dims, img has %d", out->ndim, img->ndim);
    if (nstages < 0 || nstages > B2F_MAXSTAGES) return fail(B2F_EARG, "bad stage count %d", nstages);
    if (nstages > 0 && !stages) return fail(B2F_EARG, "NULL stages");
    const int N = P.ndim = img->ndim;
    P.img_ax = axes_of(img);
    P.out_ax = axes_of(out);
    P.stages.resize(nstages);
    P.active.clear();
    int64_t sum_first[B2F_MAXDIM] = {0, 0, 0, 0}, sum_last[B2F_MAXDIM] = {0, 0, 0, 0};
    for (int s = 0; s < nstages; ++s) {
        int rc = analyse_stage(&stages[s], N, P.stages[s]);
        if (rc) return rc;
        if (!P.stages[s].copy) P.active.push_back(s);
        for (int d = 0; d < B2F_MAXDIM; ++d) {  // accumulate_padding, src/border.jl:640-642
            sum_first[d] += P.stages[s].lo[d];
            sum_last[d] += P.stages[s].hi[d];
        }
    }
    P.style = border->style;
    if (P.style < B2F_REPLICATE || P.style > B2F_NOPAD) return fail(B2F_EARG, "border style %d unrecognized", P.style);
    for (int d = 0; d < B2F_MAXDIM; ++d) P.pad_lo[d] = P.pad_hi[d] = 0;
    if (P.style <= B2F_FILL) {
        if (border->npad == 0) {  // Pad{0}(kernel): lo = max(0,-first), hi = max(0,last)  (src/border.jl:602-606)
            for (int d = 0; d < N; ++d) {
                P.pad_lo[d] = sum_first[d] < 0 ? -sum_first[d] : 0;
                P.pad_hi[d] = sum_last[d] > 0 ? sum_last[d] : 0;
            }
        } else if (border->npad == N) {
            for (int d = 0; d < N; ++d) {
                if (border->lo[d] < 0 || border->hi[d] < 0) return fail(B2F_EARG, "negative padding");
                P.pad_lo[d] = border->lo[d];
                P.pad_hi[d] = border->hi[d];
            }
        } else {
            return fail(B2F_EARG, "border lacks the proper padding sizes for an array with %d dimensions", N);
        }
    }
    P.padded_ax = P.img_ax;
    for (int d = 0; d < N; ++d) { P.padded_ax.lo[d] -= P.pad_lo[d]; P.padded_ax.hi[d] += P.pad_hi[d]; }
    P.roi = P.out_ax;
    if (roi_lo && roi_hi)
        for (int d = 0; d < N; ++d) { P.roi.lo[d] = roi_lo[d]; P.roi.hi[d] = roi_hi[d]; }

    // fill value goes through eltype(img) (src/borderarray.jl:11-20)
    P.fill = 0.0;
    if (P.style == B2F_FILL) {
        double fv = border->fill;
        if (img->dtype == B2F_N0F8) {
            double q = std::nearbyint(fv * 255.0);
            if (!(q >= 0 && q <= 255)) return fail(B2F_EARG, "fill value %g not representable as N0f8", fv);
            P.fill = (out->dtype == B2F_F32) ? (double)n0f8_to_f32((unsigned)q) : n0f8_to_f64((unsigned)q);
            if (is_int_dtype(out->dtype)) return fail(B2F_EARG, "N0f8 image with integer output");
        } else if (is_int_dtype(img->dtype)) {
            int64_t lo_ = 0, hi_ = 0;
            int_range(img->dtype, lo_, hi_);
            if (std::floor(fv) != fv || fv < (double)lo_ || fv > (double)hi_)
                return fail(B2F_EARG, "fill value %g not representable in the image eltype", fv);
            P.fill = fv;
        } else if (img->dtype == B2F_F32) {
            P.fill = (double)(float)fv;
        } else {
            P.fill = fv;
        }
        if (is_int_dtype(out->dtype)) {
            int64_t lo_ = 0, hi_ = 0;
            int_range(out->dtype, lo_, hi_);
            if (std::floor(P.fill) != P.fill || P.fill < (double)lo_ || P.fill > (double)hi_)
                return fail(B2F_EINEXACT, "fill value %g not representable in eltype(out)", P.fill);
        }
    }
    if (P.img_ax.empty() || P.roi.empty()) { P.region.clear(); return 0; }  // nothing to do

    // inds must be inbounds for out (src/imfilter.jl:604-609)
    for (int d = 0; d < N; ++d)
        if (P.roi.lo[d] < P.out_ax.lo[d] || P.roi.hi[d] > P.out_ax.hi[d])
            return fail(B2F_EDIM, "output indices disagree with requested indices (axis %d)", d);
    if (P.style == B2F_REFLECT)
        for (int d = 0; d < N; ++d)
            if ((P.pad_lo[d] > 0 || P.pad_hi[d] > 0) && P.img_ax.len(d) < 2)
                return fail(B2F_EARG, "reflect padding of a length-1 axis (DivideError in the reference)");

    // stage regions: region[a] = roi expanded by the extents of the stages after a
    const int na = (int)P.active.size();
    P.region.resize(na);
    Box cur = P.roi;
    for (int a = na - 1; a >= 0; --a) {
        P.region[a] = cur;
        const StageInfo &si = P.stages[P.active[a]];
        for (int d = 0; d < B2F_MAXDIM; ++d) { cur.lo[d] += si.lo[d]; cur.hi[d] += si.hi[d]; }
    }
    // `cur` is now the set of input indices the cascade reads: it must lie inside the padded input
    // (src/imfilter.jl:610-615), and every intermediate region inside the padded-size temporaries.
    for (int d = 0; d < N; ++d) {
        if (cur.lo[d] < P.padded_ax.lo[d] || cur.hi[d] > P.padded_ax.hi[d])
            return fail(B2F_EDIM, "requested indices and kernel indices do not agree with indices of padded input (axis %d)", d);
        for (int a = 0; a + 1 < na; ++a)
            if (P.region[a].lo[d] < P.padded_ax.lo[d] || P.region[a].hi[d] > P.padded_ax.hi[d])
                return fail(B2F_EDIM, "stage region exceeds the temporary buffer (axis %d)", d);
    }
    return 0;
}

// ---- per-thread context: device scratch for host-mode calls ------------------------------------------
struct Ctx {
    bool pool_ready = false;
    int device = -1;
};
static thread_local Ctx g_ctx;

int ensure_ctx() {
    int dev = 0;
    B2F_CUDA(cudaGetDevice(&dev));
    if (!g_ctx.pool_ready || g_ctx.device != dev) {
        cudaMemPool_t pool;
        B2F_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t thr = UINT64_MAX;  // keep freed blocks cached: no re-allocation cost per call
        B2F_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        g_ctx.pool_ready = true;
        g_ctx.device = dev;
    }
    return 0;
}

int stage_in(const b2f_array *a, Staged &s, cudaStream_t st, bool copy) {
    int64_t n = 1;
    for (int d = 0; d < a->ndim; ++d) n *= a->dims[d] < 0 ? 0 : a->dims[d];
    s.bytes = (size_t)n * dtype_size(a->dtype);
    if (a->mem == B2F_DEVICE) { s.dptr = a->ptr; s.owned = false; return 0; }
    if (s.bytes == 0) { s.dptr = nullptr; return 0; }
    if (!a->ptr) return fail(B2F_EARG, "NULL array pointer");
    B2F_CUDA(cudaMallocAsync(&s.dptr, s.bytes, st));
    s.owned = true;
    if (copy) B2F_CUDA(cudaMemcpyAsync(s.dptr, a->ptr, s.bytes, cudaMemcpyHostToDevice, st));
    return 0;
}
void release(Staged &s, cudaStream_t st) {
    if (s.owned && s.dptr) cudaFreeAsync(s.dptr, st);
    s.dptr = nullptr;
    s.owned = false;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

const char *b2f_version(void) { return "b2f 0.1 (CUDA sm_100a)"; }
const char *b2f_last_error(void) { return g_err.c_str(); }
int b2f_is_device_library(void) { return 1; }
int64_t b2f_launch_count(void) { return g_launches; }
void b2f_reset_launch_count(void) { g_launches = 0; }
const char *b2f_last_path(void) { return g_path.c_str(); }
int b2f_set_accum_mode(int32_t mode) {
    if (mode != B2F_ACCUM_EXACT && mode != B2F_ACCUM_FMA) return fail(B2F_EARG, "unknown accumulate mode %d", (int)mode);
    const int prev = g_accum;
    g_accum = mode;
    return prev;
}

int b2f_set_device(int device) { B2F_CUDA(cudaSetDevice(device)); return 0; }
int b2f_device_count(int *count) {
    if (!count) return fail(B2F_EARG, "NULL argument");
    B2F_CUDA(cudaGetDeviceCount(count));
    return 0;
}
int b2f_sm_count(int *count) {
    if (!count) return fail(B2F_EARG, "NULL argument");
    *count = sm_count();
    return 0;
}
int b2f_malloc(void **dptr, uint64_t bytes) { B2F_CUDA(cudaMalloc(dptr, bytes)); return 0; }
int b2f_free(void *dptr) { B2F_CUDA(cudaFree(dptr)); return 0; }
int b2f_host_alloc(void **hptr, uint64_t bytes) { B2F_CUDA(cudaHostAlloc(hptr, bytes, cudaHostAllocDefault)); return 0; }
int b2f_host_free(void *hptr) { B2F_CUDA(cudaFreeHost(hptr)); return 0; }
int b2f_host_register(void *hptr, uint64_t bytes) {
    if (!hptr) return fail(B2F_EARG, "NULL argument");
    B2F_CUDA(cudaHostRegister(hptr, bytes, cudaHostRegisterDefault));
    return 0;
}
int b2f_host_unregister(void *hptr) {
    if (!hptr) return fail(B2F_EARG, "NULL argument");
    B2F_CUDA(cudaHostUnregister(hptr));
    return 0;
}
int b2f_memcpy_h2d(void *dptr, const void *hptr, uint64_t bytes) {
    B2F_CUDA(cudaMemcpy(dptr, hptr, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int b2f_memcpy_d2h(void *hptr, const void *dptr, uint64_t bytes) {
    B2F_CUDA(cudaMemcpy(hptr, dptr, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int b2f_sync(void) { B2F_CUDA(cudaDeviceSynchronize()); return 0; }

// ---- peer mapping of device memory across processes (one process per GPU; the NVLink halo path) ------------------
// cudaIpcGetMemHandle names the whole ALLOCATION that contains dptr, so the offset of dptr inside it travels along
// (cuMemGetAddressRange, reached through the runtime so the library has no link-time dependency on libcuda).
int b2f_ipc_export(const void *dptr, void *handle64, uint64_t *offset) {
    if (!dptr || !handle64 || !offset) return fail(B2F_EARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    B2F_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return fail(B2F_ECUDA, "cuMemGetAddressRange is unavailable");
    unsigned long long base = 0;
    size_t size = 0;
    if (((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)dptr) != 0)
        return fail(B2F_ECUDA, "cuMemGetAddressRange failed (not a device allocation?)");
    cudaIpcMemHandle_t h;
    B2F_CUDA(cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base));
    memcpy(handle64, &h, 64);
    *offset = (uint64_t)((uintptr_t)dptr - (uintptr_t)base);
    return 0;
}
int b2f_ipc_open(const void *handle64, uint64_t offset, void **dptr) {
    if (!handle64 || !dptr) return fail(B2F_EARG, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *base = nullptr;
    B2F_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dptr = (char *)base + offset;
    return 0;
}
int b2f_ipc_close(void *dptr, uint64_t offset) {
    if (!dptr) return 0;
    B2F_CUDA(cudaIpcCloseMemHandle((char *)dptr - offset));
    return 0;
}

static int check_types(const b2f_array *img, const b2f_array *out, const Plan &P) {
    if (img->dtype < B2F_U8 || img->dtype > B2F_U32) return fail(B2F_EARG, "unsupported image dtype %d", img->dtype);
    if (out->dtype < B2F_U8 || out->dtype > B2F_U32 || out->dtype == B2F_N0F8)
        return fail(B2F_EARG, "unsupported output dtype %d", out->dtype);
    if (is_int_dtype(out->dtype)) {
        if (img->dtype == B2F_N0F8) return fail(B2F_EARG, "N0f8 image with integer output");
        if (!is_int_dtype(img->dtype))
            return fail(B2F_ENOTSUP, "integer output from a floating-point image is not accelerated");
        for (int a : P.active)
            if (P.stages[a].s->tap_dtype != B2F_TAPS_INT)
                return fail(B2F_ENOTSUP, "integer output with non-integer taps is not accelerated");
    }
    return 0;
}

// ---- pipelined host path -------------------------------------------------------------------------------------------------------
// A large host-to-host call is PCIe-bound (C5: 4.3 GB up, 3 ms of kernel, 4.3 GB down).  When both arrays are pinned
// (b2f_host_alloc / b2f_host_register) and the cascade has a slab form (b2f_imfilter_slab: separable, whole-array output), the
// array is cut into chunks of planes along its last axis and the three phases run as a pipeline on three streams — upload of
// chunk c+1, kernel of chunk c (its halo planes are its neighbours' planes in the same device buffer, the array's faces are
// handled in global coordinates by the slab form), download of chunk c-1 — so that both directions of the link are busy at
// once.  Results are those of the unpipelined call (the slab form is tested against the whole-array path plane by plane).
struct HostPipe {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev;
    int device = -1;
};
static thread_local HostPipe g_pipe;

static bool host_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// returns 1 when the call was handled here, 0 when the ordinary path should run, < 0 on error
static int imfilter_host_pipelined(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int nstages,
                                   const b2f_border *border, const Plan &P, cudaStream_t st) {
    static const bool enabled = !(getenv("B2F_HOST_PIPELINE") && atoi(getenv("B2F_HOST_PIPELINE")) == 0);
    if (!enabled || img->mem != B2F_HOST || out->mem != B2F_HOST || img->ndim < 2 || getenv("B2F_FORCE_PATH")) return 0;
    const int N = img->ndim, last = N - 1;
    for (int d = 0; d < N; ++d)
        if (img->dims[d] != out->dims[d] || img->origin[d] != out->origin[d]) return 0;
    if (P.style > B2F_FILL || !sepnd_applicable(P, img->dtype, out->dtype)) return 0;
    int64_t zlo = 0, zhi = 0;
    for (int a : P.active) { zlo += P.stages[a].lo[last]; zhi += P.stages[a].hi[last]; }
    const int64_t h_lo = zlo < 0 ? -zlo : 0, h_hi = zhi > 0 ? zhi : 0, h = h_lo > h_hi ? h_lo : h_hi;
    if (P.style == B2F_CIRCULAR && h > 0) return 0;            // the wrap-around halo of the first chunk is the last one uploaded
    int64_t plane = 1;
    for (int d = 0; d < last; ++d) plane *= img->dims[d];
    const int64_t Z = img->dims[last];
    const size_t in_pb = (size_t)plane * dtype_size(img->dtype), out_pb = (size_t)plane * dtype_size(out->dtype);
    if ((in_pb + out_pb) * (size_t)Z < ((size_t)64 << 20)) return 0;          // small calls: one upload, one launch, one download
    int64_t nchunk = (int64_t)(((in_pb + out_pb) * (size_t)Z + ((size_t)32 << 20) - 1) / ((size_t)32 << 20));   // ~32 MiB per chunk
    if (nchunk > 64) nchunk = 64;
    while (nchunk > 1 && Z / nchunk < 2 * h + 1) --nchunk;
    if (nchunk < 3) return 0;
    if (!host_pinned(img->ptr) || !host_pinned(out->ptr)) return 0;      // pageable memory is staged synchronously: no overlap to win
    int dev = 0;
    B2F_CUDA(cudaGetDevice(&dev));
    HostPipe &hp = g_pipe;
    if (hp.device != dev) {
        hp = HostPipe();
        B2F_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
        B2F_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
        hp.ev.resize(2 * 64 + 2);
        for (auto &e : hp.ev) B2F_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        hp.device = dev;
    }
    cudaEvent_t *ev_in = hp.ev.data(), *ev_k = hp.ev.data() + 64, ev_a = hp.ev[128], ev_done = hp.ev[129];
    char *d_in = nullptr, *d_out = nullptr;
    B2F_CUDA(cudaMallocAsync((void **)&d_in, in_pb * (size_t)Z, st));
    AsyncFrees guard(st);
    guard.push_back(d_in);
    B2F_CUDA(cudaMallocAsync((void **)&d_out, out_pb * (size_t)Z, st));
    guard.push_back(d_out);
    const int64_t base = Z / nchunk, extra = Z % nchunk;
    auto first_of = [&](int64_t c) { return c * base + (c < extra ? c : extra); };
    // everything that queues work on the side streams runs inside this lambda, so that EVERY exit — error returns included —
    // passes through the join below before the buffers are released
    auto run = [&]() -> int {
        B2F_CUDA(cudaEventRecord(ev_a, st));
        B2F_CUDA(cudaStreamWaitEvent(hp.s_in, ev_a, 0));
        B2F_CUDA(cudaStreamWaitEvent(hp.s_out, ev_a, 0));
        for (int64_t c = 0; c < nchunk; ++c) {
            const int64_t z0 = first_of(c), z1 = first_of(c + 1);
            B2F_CUDA(cudaMemcpyAsync(d_in + (size_t)z0 * in_pb, (const char *)img->ptr + (size_t)z0 * in_pb, (size_t)(z1 - z0) * in_pb,
                                     cudaMemcpyHostToDevice, hp.s_in));
            B2F_CUDA(cudaEventRecord(ev_in[c], hp.s_in));
        }
        for (int64_t c = 0; c < nchunk; ++c) {
            const int64_t z0 = first_of(c), z1 = first_of(c + 1);
            B2F_CUDA(cudaStreamWaitEvent(st, ev_in[c + 1 < nchunk ? c + 1 : c], 0));      // the upper halo lives in the next chunk
            b2f_array a = *img, o = *out;
            a.mem = o.mem = B2F_DEVICE;
            a.ptr = d_in + (size_t)z0 * in_pb;
            o.ptr = d_out + (size_t)z0 * out_pb;
            a.dims[last] = o.dims[last] = z1 - z0;
            const int64_t nlo = z0 < h_lo ? z0 : h_lo, nhi = Z - z1 < h_hi ? Z - z1 : h_hi;
            int rc = b2f_imfilter_slab(&a, &o, stages, nstages, border, Z, z0, nlo ? d_in + (size_t)(z0 - nlo) * in_pb : nullptr, nlo,
                                       nhi ? d_in + (size_t)z1 * in_pb : nullptr, nhi, st);
            if (rc) return rc;
            B2F_CUDA(cudaEventRecord(ev_k[c], st));
            B2F_CUDA(cudaStreamWaitEvent(hp.s_out, ev_k[c], 0));
            B2F_CUDA(cudaMemcpyAsync((char *)out->ptr + (size_t)z0 * out_pb, d_out + (size_t)z0 * out_pb, (size_t)(z1 - z0) * out_pb,
                                     cudaMemcpyDeviceToHost, hp.s_out));
        }
        return 0;
    };
    const int rc = run();
    // join: nothing may outlive the call (the buffers are freed on st, the caller owns the host arrays again)
    cudaStreamSynchronize(hp.s_in);
    cudaStreamSynchronize(hp.s_out);
    cudaError_t e = cudaStreamSynchronize(st);
    (void)ev_done;
    if (rc) return rc;
    if (e != cudaSuccess) return fail(B2F_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    return 1;
}

static int imfilter_planes(const b2f_array *img, const b2f_array *outs, int nplanes, const b2f_stage *stages,
                           int nstages_each, const b2f_border *border, const int64_t *roi_lo,
                           const int64_t *roi_hi, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (nplanes < 1 || nplanes > 4) return fail(B2F_EARG, "bad plane count %d", nplanes);
    std::vector<Plan> plans(nplanes);
    bool nothing = false;
    for (int p = 0; p < nplanes; ++p) {
        int rc = make_plan(img, &outs[p], stages + (size_t)p * nstages_each, nstages_each, border, roi_lo, roi_hi, plans[p]);
        if (rc) return rc;
        rc = check_types(img, &outs[p], plans[p]);
        if (rc) return rc;
        if (plans[p].img_ax.empty() || plans[p].roi.empty()) nothing = true;
    }
    if (nothing) { set_path("empty"); return 0; }
    int rc = ensure_ctx();
    if (rc) return rc;
    if (nplanes == 1 && !roi_lo && !roi_hi) {
        rc = imfilter_host_pipelined(img, &outs[0], stages, nstages_each, border, plans[0], st);
        if (rc < 0) return rc;
        if (rc == 1) return 0;
    }

    bool any_host = img->mem == B2F_HOST;
    for (int p = 0; p < nplanes; ++p) any_host = any_host || outs[p].mem == B2F_HOST;
    Staged sin;
    std::vector<Staged> souts(nplanes);
    rc = stage_in(img, sin, st, true);
    // a partial roi leaves the rest of a host `out` untouched, so it has to be uploaded first
    for (int p = 0; p < nplanes && !rc; ++p) rc = stage_in(&outs[p], souts[p], st, roi_lo != nullptr);
    if (!rc) {
        std::vector<void *> dout(nplanes);
        std::vector<int> odt(nplanes);
        for (int p = 0; p < nplanes; ++p) { dout[p] = souts[p].dptr; odt[p] = outs[p].dtype; }
        const char *force = getenv("B2F_FORCE_PATH");   // debugging knob: "fused2d" | "sepnd" | "generic"
        const bool allow_stream = !force, allow_fused = !force || !strcmp(force, "fused2d");
        // 18 .. 32 taps: the chunked-tap passes of sepnd / longtap beat the shared-memory tiled fused2d (25 x 25 taps on
        // 8192^2 Float32: 0.45 vs 1.15 ms, profiles/r2_longtap.jsonl) wherever sepnd applies (whole-array outputs)
        bool long_taps = false;
        for (int p = 0; p < nplanes; ++p)
            for (int a : plans[p].active) {
                const StageInfo &si = plans[p].stages[a];
                if (si.s->kind == B2F_STAGE_1D && si.s->len[si.s->axis] > 17) long_taps = true;
            }
        const bool prefer_sepnd = !force && long_taps && nplanes == 1 && sepnd_applicable(plans[0], img->dtype, odt[0]);
        if (allow_stream && stream2d_applicable(plans.data(), nplanes, img->dtype, odt.data())) {
            rc = run_stream2d(plans.data(), nplanes, sin.dptr, img->dtype, dout.data(), odt.data(), st);
        } else if (!prefer_sepnd && allow_fused && fused2d_applicable(plans.data(), nplanes, img->dtype, odt.data())) {
            rc = run_fused2d(plans.data(), nplanes, sin.dptr, img->dtype, dout.data(), odt.data(), st);
        } else {
            for (int p = 0; p < nplanes && !rc; ++p) {
                if (!force && stream3d_applicable(plans[p], img->dtype, odt[p]))
                    rc = run_stream3d(plans[p], sin.dptr, dout[p], st);
                else if ((!force || !strcmp(force, "sepnd")) && sepnd_applicable(plans[p], img->dtype, odt[p]))
                    rc = run_sepnd(plans[p], sin.dptr, img->dtype, dout[p], odt[p], st);
                else if (!force && dense2d_applicable(plans[p], img->dtype, odt[p]))
                    rc = run_dense2d(plans[p], sin.dptr, img->dtype, dout[p], odt[p], st);
                else if (!force && dense3d_applicable(plans[p], img->dtype, odt[p]))
                    rc = run_dense3d(plans[p], sin.dptr, img->dtype, dout[p], odt[p], st);
                else
                    rc = run_generic(plans[p], sin.dptr, img->dtype, dout[p], odt[p], st);
            }
        }
        for (int p = 0; p < nplanes && !rc; ++p)
            if (outs[p].mem == B2F_HOST && souts[p].bytes) {
                cudaError_t e = cudaMemcpyAsync(outs[p].ptr, souts[p].dptr, souts[p].bytes, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess) rc = fail(B2F_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
            }
    }
    release(sin, st);
    for (auto &s : souts) release(s, st);
    if (any_host) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && !rc) rc = fail(B2F_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

int b2f_imfilter(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                 const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi, void *stream) {
    if (!img || !out) return fail(B2F_EARG, "NULL argument");
    return imfilter_planes(img, out, 1, stages, nstages, border, roi_lo, roi_hi, stream);
}

int b2f_imgradients(const b2f_array *img, const b2f_array *outs, int32_t nplanes, const b2f_stage *stages,
                    int32_t nstages_each, const b2f_border *border, void *stream) {
    if (!img || !outs) return fail(B2F_EARG, "NULL argument");
    return imfilter_planes(img, outs, nplanes, stages, nstages_each, border, nullptr, nullptr, stream);
}

int b2f_mapwindow_extrema(const b2f_array *img, const b2f_array *out_min, const b2f_array *out_max,
                          int32_t interleaved, const int64_t *win_lo, const int64_t *win_hi,
                          const b2f_border *border, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !border || !win_lo || !win_hi || (!out_min && !out_max)) return fail(B2F_EARG, "NULL argument");
    if (interleaved && !out_min) return fail(B2F_EARG, "interleaved output needs out_min");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM) return fail(B2F_ENOTSUP, "ndim %d not supported", img->ndim);
    const b2f_array *ref = out_min ? out_min : out_max;
    if (ref->dtype != img->dtype || (out_max && out_max->dtype != img->dtype) || ref->ndim != img->ndim)
        return fail(B2F_EARG, "extrema outputs must have the image eltype and rank");
    const int N = img->ndim;
    if (border->style == B2F_NOPAD) return fail(B2F_ENOTSUP, "NoPad() is not supported by mapwindow");
    for (int d = 0; d < N; ++d)
        if (win_lo[d] > win_hi[d]) return fail(B2F_EARG, "empty window");
    Box ia = axes_of(img), oa = axes_of(ref);
    if (out_min && out_max && !interleaved) {
        Box ob = axes_of(out_max);
        for (int d = 0; d < N; ++d)
            if (ob.lo[d] != oa.lo[d] || ob.hi[d] != oa.hi[d]) return fail(B2F_EDIM, "out_min and out_max axes differ");
    }
    if (border->style == B2F_INNER)
        for (int d = 0; d < N; ++d)
            if (win_hi[d] - win_lo[d] + 1 > ia.len(d))
                return fail(B2F_EDIM, "window is larger than the image along axis %d: no interior for Inner()", d);
    if (ia.empty() || oa.empty()) { set_path("empty"); return 0; }
    for (int d = 0; d < N; ++d) {
        if (oa.lo[d] < ia.lo[d] || oa.hi[d] > ia.hi[d]) return fail(B2F_EDIM, "output axes exceed image axes");
        if (border->style == B2F_INNER && (oa.lo[d] + win_lo[d] < ia.lo[d] || oa.hi[d] + win_hi[d] > ia.hi[d]))
            return fail(B2F_EDIM, "output axes are not in the interior for Inner()");
    }
    double fill = border->fill;
    if (border->style == B2F_FILL) {
        if (is_int_dtype(img->dtype) || img->dtype == B2F_N0F8) {
            int64_t lo_ = 0, hi_ = 255;
            double fv = fill;
            if (img->dtype == B2F_N0F8) fv = std::nearbyint(fill * 255.0); else int_range(img->dtype, lo_, hi_);
            if (std::floor(fv) != fv || fv < (double)lo_ || fv > (double)hi_)
                return fail(B2F_EARG, "fill value %g not representable in eltype(img)", fill);
            fill = fv;
        } else if (img->dtype == B2F_F32) {
            fill = (double)(float)fill;
        }
    }
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin, smn, smx;
    rc = stage_in(img, sin, st, true);
    b2f_array pair_desc;
    if (!rc && out_min) {
        pair_desc = *out_min;
        if (interleaved) { pair_desc.dims[0] *= 2; }
        rc = stage_in(&pair_desc, smn, st, false);
    }
    if (!rc && out_max && !interleaved) rc = stage_in(out_max, smx, st, false);
    if (!rc)
        rc = run_extrema(img, sin.dptr, smn.dptr, interleaved ? nullptr : smx.dptr, interleaved, oa, win_lo, win_hi,
                         border->style, fill, st);
    bool any_host = img->mem == B2F_HOST;
    if (!rc && out_min && out_min->mem == B2F_HOST && smn.bytes) {
        any_host = true;
        cudaError_t e = cudaMemcpyAsync(out_min->ptr, smn.dptr, smn.bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = fail(B2F_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
    }
    if (!rc && out_max && !interleaved && out_max->mem == B2F_HOST && smx.bytes) {
        any_host = true;
        cudaError_t e = cudaMemcpyAsync(out_max->ptr, smx.dptr, smx.bytes, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = fail(B2F_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
    }
    release(sin, st);
    release(smn, st);
    release(smx, st);
    if (any_host) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && !rc) rc = fail(B2F_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

}  // extern "C"
