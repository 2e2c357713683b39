// points.cu — the consumers of the LoG path that SURVEY §8(f) ranks next: local-extrema scan and the blob_LoG plumbing.
//
//   b2f_findlocalextrema   findlocalmaxima / findlocalminima      reference src/extrema.jl:107-164
//   b2f_scale_into_slice   multiLoG's slice write and `.*= -σ`     reference src/extrema.jl:94-105
//   b2f_maxabs             maximum(abs, img)                      reference src/extrema.jl:85
//   b2f_gather             img_LoG[x] at the peaks                reference src/extrema.jl:86-90
//   b2f_na_prepare / b2f_divide / b2f_normalize_dims   the element-wise pieces of the NA() border (§8f rank 2)
//                                                                 reference src/imfilter.jl:282-318, 1110-1127, 1234-1250
//
// All of it is data-parallel compare / copy work: one thread per element, coalesced along the first axis, HBM-bound.
// The peak list comes back in the reference's order (column-major ascending): per-block counts, one scan, an ordered
// scatter (warp ballots give the rank of every peak inside its block).
#include <cstring>

#include "common.cuh"

namespace b2f {

struct PtGeom {
    int ndim;
    long long dims[B2F_MAXDIM];
    int half[B2F_MAXDIM];      // window >> 1
    int clip[B2F_MAXDIM];      // 1: first and last index along the axis are not eligible (edges[d] == false)
    long long n;
};

constexpr int PT_BLOCK = 256;

template <typename T> __device__ __forceinline__ T pt_load(const void *p, int dt, long long i);
template <> __device__ __forceinline__ double pt_load<double>(const void *p, int dt, long long i) {
    return dt == B2F_N0F8 ? (double)((const uint8_t *)p)[i] : load_elem<double>(p, dt, i);   // raw codes order like values
}
template <> __device__ __forceinline__ long long pt_load<long long>(const void *p, int, long long i) {
    return ((const long long *)p)[i];
}

// isextrema of src/extrema.jl:137-160: strict comparison against every neighbour of the window that lies inside the
// array; NaN compares false, so a NaN is never a peak and never lets a neighbour be one
template <typename T>
__device__ __forceinline__ bool pt_is_extremum(const void *img, int dt, const PtGeom &G, long long lin, bool minima) {
    long long c[B2F_MAXDIM], r = lin;
#pragma unroll
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        c[d] = r % G.dims[d];
        r /= G.dims[d];
        if (G.clip[d] && (c[d] == 0 || c[d] == G.dims[d] - 1)) return false;
    }
    const T v = pt_load<T>(img, dt, lin);
    long long stride[B2F_MAXDIM];
    stride[0] = 1;
#pragma unroll
    for (int d = 1; d < B2F_MAXDIM; ++d) stride[d] = stride[d - 1] * G.dims[d - 1];
    for (int j3 = -G.half[3]; j3 <= G.half[3]; ++j3) {
        if ((unsigned long long)(c[3] + j3) >= (unsigned long long)G.dims[3]) continue;
        for (int j2 = -G.half[2]; j2 <= G.half[2]; ++j2) {
            if ((unsigned long long)(c[2] + j2) >= (unsigned long long)G.dims[2]) continue;
            for (int j1 = -G.half[1]; j1 <= G.half[1]; ++j1) {
                if ((unsigned long long)(c[1] + j1) >= (unsigned long long)G.dims[1]) continue;
                const long long base = lin + j3 * stride[3] + j2 * stride[2] + j1 * stride[1];
                for (int j0 = -G.half[0]; j0 <= G.half[0]; ++j0) {
                    if ((unsigned long long)(c[0] + j0) >= (unsigned long long)G.dims[0]) continue;
                    if (j0 == 0 && j1 == 0 && j2 == 0 && j3 == 0) continue;
                    const T w = pt_load<T>(img, dt, base + j0);
                    if (!(minima ? v < w : v > w)) return false;
                }
            }
        }
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(PT_BLOCK) pt_flag_kernel(const void *img, int dt, PtGeom G, int minima,
                                                           unsigned char *flags, unsigned *block_counts) {
    const long long lin = (long long)blockIdx.x * PT_BLOCK + threadIdx.x;
    const bool f = lin < G.n && pt_is_extremum<T>(img, dt, G, lin, minima != 0);
    if (lin < G.n) flags[lin] = f ? 1 : 0;
    const int cnt = __syncthreads_count(f);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned)cnt;
}

// exclusive scan of the per-block counts (one CTA; every thread owns a contiguous chunk); total -> *total
__global__ void __launch_bounds__(1024) pt_scan_kernel(const unsigned *counts, long long *offsets, long long nblocks,
                                                       long long *total) {
    __shared__ long long part[1024];
    const long long chunk = (nblocks + 1023) / 1024;
    const long long b0 = min((long long)threadIdx.x * chunk, nblocks), b1 = min(b0 + chunk, nblocks);
    long long s = 0;
    for (long long b = b0; b < b1; ++b) s += counts[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) { const long long v = part[t]; part[t] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long b = b0; b < b1; ++b) { offsets[b] = run; run += counts[b]; }
}

__global__ void __launch_bounds__(PT_BLOCK) pt_scatter_kernel(const unsigned char *flags, const long long *offsets, long long n,
                                                              long long *idx, long long cap) {
    __shared__ unsigned wsum[PT_BLOCK / 32];
    const long long lin = (long long)blockIdx.x * PT_BLOCK + threadIdx.x;
    const bool f = lin < n && flags[lin];
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    unsigned before = 0;
    for (int w = 0; w < warp; ++w) before += wsum[w];
    if (f) {
        const long long pos = offsets[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1));
        if (pos < cap) idx[pos] = lin;
    }
}

__global__ void pt_scale_slice_kernel(const void *src, int sdt, void *stack, int odt, long long n, long long S, long long slice,
                                      double scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = load_elem<double>(src, sdt, i) * scale;      // slice .*= -σ: product in Float64, stored as eltype
        if (odt == B2F_F32) ((float *)stack)[slice + S * i] = (float)v;
        else ((double *)stack)[slice + S * i] = v;
    }
}

__global__ void pt_maxabs_kernel(const void *img, int dt, long long n, unsigned long long *bits, int *nan_seen) {
    double m = 0.0;
    bool nan = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = fabs(load_elem<double>(img, dt, i));
        if (v != v) nan = true; else m = fmax(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (__any_sync(0xffffffffu, nan) && (threadIdx.x & 31) == 0) atomicOr(nan_seen, 1);
    if ((threadIdx.x & 31) == 0) atomicMax(bits, (unsigned long long)__double_as_longlong(m));   // m >= 0: bits order like values
}

__global__ void pt_gather_kernel(const void *arr, int dt, const long long *idx, long long n, double *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = load_elem<double>(arr, dt, idx[i]);
}

// ---- NA() border (reference src/imfilter.jl:282-318, 1110-1127, 1234-1250): the element-wise pieces around the two FIR calls ----
__device__ __forceinline__ bool pt_is_na(double v, int mode) {      // 0: isnan, 1: !isfinite, 2: never
    return mode == 0 ? (v != v) : mode == 1 ? !(fabs(v) <= 1.79769313486231570815e308) : false;
}
__global__ void pt_na_prepare_kernel(const void *img, int dt, long long n, int mode, void *imgtmp, int tdt, void *valid, int vdt,
                                     int *hasna) {
    bool any = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = load_elem<double>(img, dt, i);
        const bool na = pt_is_na(v, mode);
        any = any || na;
        if (imgtmp) {                                   // imgtmp[naflag] .= zero(T)
            if (tdt == B2F_F32) ((float *)imgtmp)[i] = na ? 0.f : (dt == B2F_N0F8 ? n0f8_to_f32(((const uint8_t *)img)[i]) : (float)v);
            else ((double *)imgtmp)[i] = na ? 0.0 : v;
        }
        if (valid) {                                    // validpixels = !naflag
            if (vdt == B2F_F32) ((float *)valid)[i] = na ? 0.f : 1.f;
            else ((double *)valid)[i] = na ? 0.0 : 1.0;
        }
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) atomicOr(hasna, 1);
}
__global__ void pt_divide_kernel(void *out, int odt, const void *den, int ddt, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (odt == B2F_F32 && ddt == B2F_F32) ((float *)out)[i] = __fdiv_rn(((float *)out)[i], ((const float *)den)[i]);
        else {                                          // promoted to Float64, stored as eltype(out)
            const double q = __ddiv_rn(load_elem<double>(out, odt, i), load_elem<double>(den, ddt, i));
            if (odt == B2F_F32) ((float *)out)[i] = (float)q; else ((double *)out)[i] = q;
        }
    }
}
struct PtFactors { const double *f[B2F_MAXDIM]; long long dims[B2F_MAXDIM]; int ndim; };
__global__ void pt_normalize_dims_kernel(void *out, int odt, long long n, PtFactors F) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        double t = load_elem<double>(out, odt, i);      // tmp = A[I] / f1[I1]; tmp /= f2[I2]; ...  (src/imfilter.jl:1241-1250)
#pragma unroll
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            if (d < F.ndim) {
                const long long c = r % F.dims[d];
                r /= F.dims[d];
                t = __ddiv_rn(t, F.f[d][c]);
            }
        }
        if (odt == B2F_F32) ((float *)out)[i] = (float)t; else ((double *)out)[i] = t;
    }
}

// ---- mapwindow(median!, img, window) (reference src/mapwindow.jl:270-333 generic path with
// f = median!, Statistics.median!: NaN if any NaN, middle of the sorted window, x/2 + y/2 for even lengths) ----------------
constexpr int PT_MAXWIN = 128;
struct PtWin {
    int ndim;
    long long dims[B2F_MAXDIM];       // image extent
    long long odims[B2F_MAXDIM];      // output extent
    long long ooff[B2F_MAXDIM];       // image coordinate (0-based) of output element 0
    long long ostep[B2F_MAXDIM];      // image-coordinate step between output elements (`indices=` stride)
    int wlo[B2F_MAXDIM], wn[B2F_MAXDIM];
    int style;
    double fill;
    long long nout;
    int wtotal;
};

// window position k (0-based image coordinate, possibly outside) along an axis of length n, for the window [a, b]:
// copy_win! pads the window's in-image part `inner` = [max(a,0), min(b,n-1)] by padindex (src/mapwindow.jl:310-317,
// src/border.jl:564-590): the border remap is taken relative to `inner`, not to the whole image.  -1: Fill value.
__host__ __device__ inline long long pt_win_index(int style, long long k, long long a, long long b, long long n) {
    if (style == B2F_FILL) return (k >= 0 && k < n) ? k : -1;
    const long long lo = a > 0 ? a : 0, hi = b < n - 1 ? b : n - 1, len = hi - lo + 1;
    if (len == 1) return lo;
    return lo + remap_index(style, k - lo, len);
}

template <typename T>
__global__ void __launch_bounds__(128) pt_median_kernel(const void *img, int dt, void *out, int odt, PtWin G) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= G.nout) return;
    long long c[B2F_MAXDIM], r = o;
#pragma unroll
    for (int d = 0; d < B2F_MAXDIM; ++d) { c[d] = (r % G.odims[d]) * G.ostep[d] + G.ooff[d]; r /= G.odims[d]; }
    T buf[PT_MAXWIN];
    bool nan = false;
    int n = 0;
    long long stride[B2F_MAXDIM];
    stride[0] = 1;
#pragma unroll
    for (int d = 1; d < B2F_MAXDIM; ++d) stride[d] = stride[d - 1] * G.dims[d - 1];
    for (int j3 = 0; j3 < G.wn[3]; ++j3) {
        const long long i3 = pt_win_index(G.style, c[3] + G.wlo[3] + j3, c[3] + G.wlo[3], c[3] + G.wlo[3] + G.wn[3] - 1, G.dims[3]);
        for (int j2 = 0; j2 < G.wn[2]; ++j2) {
            const long long i2 = pt_win_index(G.style, c[2] + G.wlo[2] + j2, c[2] + G.wlo[2], c[2] + G.wlo[2] + G.wn[2] - 1, G.dims[2]);
            for (int j1 = 0; j1 < G.wn[1]; ++j1) {
                const long long i1 = pt_win_index(G.style, c[1] + G.wlo[1] + j1, c[1] + G.wlo[1], c[1] + G.wlo[1] + G.wn[1] - 1, G.dims[1]);
                for (int j0 = 0; j0 < G.wn[0]; ++j0) {
                    const long long i0 = pt_win_index(G.style, c[0] + G.wlo[0] + j0, c[0] + G.wlo[0], c[0] + G.wlo[0] + G.wn[0] - 1, G.dims[0]);
                    T v;
                    if (i0 < 0 || i1 < 0 || i2 < 0 || i3 < 0) v = (T)G.fill;
                    else v = pt_load<T>(img, dt, i0 + i1 * stride[1] + i2 * stride[2] + i3 * stride[3]);
                    nan = nan || (v != v);
                    buf[n++] = v;
                }
            }
        }
    }
    // partial selection sort up to the upper middle element (windows are small)
    const int hiidx = n / 2;                      // 0-based: odd n -> the middle; even n -> the upper of the two
    for (int i = 0; i <= hiidx; ++i) {
        int m = i;
        for (int j = i + 1; j < n; ++j)
            if (buf[j] < buf[m]) m = j;
        const T t = buf[i]; buf[i] = buf[m]; buf[m] = t;
    }
    if (odt == B2F_F32) {
        float res;
        if (nan) res = __int_as_float(0x7fc00000);
        else if (n & 1) res = (float)buf[hiidx];
        else res = __fadd_rn(__fmul_rn((float)buf[hiidx - 1], 0.5f), __fmul_rn((float)buf[hiidx], 0.5f));
        ((float *)out)[o] = res;
    } else {
        double res;
        if (nan) res = __longlong_as_double(0x7ff8000000000000LL);
        else if (n & 1) res = (double)buf[hiidx];
        else res = __dadd_rn(__dmul_rn((double)buf[hiidx - 1], 0.5), __dmul_rn((double)buf[hiidx], 0.5));
        ((double *)out)[o] = res;
    }
}

// mapwindow(mean | sum | minimum | maximum, ...): the window is reduced on the fly in its memory (column-major) order; ACC is
// the accumulator the reference's reduction has (Float32 / Float64 / Int)
template <typename ACC>
__global__ void __launch_bounds__(128) pt_winreduce_kernel(const void *img, int dt, void *out, int odt, int op, PtWin G) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= G.nout) return;
    long long c[B2F_MAXDIM], r = o;
#pragma unroll
    for (int d = 0; d < B2F_MAXDIM; ++d) { c[d] = (r % G.odims[d]) * G.ostep[d] + G.ooff[d]; r /= G.odims[d]; }
    long long stride[B2F_MAXDIM];
    stride[0] = 1;
#pragma unroll
    for (int d = 1; d < B2F_MAXDIM; ++d) stride[d] = stride[d - 1] * G.dims[d - 1];
    ACC acc = ACC(0);
    bool first = true;
    for (int j3 = 0; j3 < G.wn[3]; ++j3) {
        const long long i3 = pt_win_index(G.style, c[3] + G.wlo[3] + j3, c[3] + G.wlo[3], c[3] + G.wlo[3] + G.wn[3] - 1, G.dims[3]);
        for (int j2 = 0; j2 < G.wn[2]; ++j2) {
            const long long i2 = pt_win_index(G.style, c[2] + G.wlo[2] + j2, c[2] + G.wlo[2], c[2] + G.wlo[2] + G.wn[2] - 1, G.dims[2]);
            for (int j1 = 0; j1 < G.wn[1]; ++j1) {
                const long long i1 = pt_win_index(G.style, c[1] + G.wlo[1] + j1, c[1] + G.wlo[1], c[1] + G.wlo[1] + G.wn[1] - 1, G.dims[1]);
                for (int j0 = 0; j0 < G.wn[0]; ++j0) {
                    const long long i0 = pt_win_index(G.style, c[0] + G.wlo[0] + j0, c[0] + G.wlo[0], c[0] + G.wlo[0] + G.wn[0] - 1, G.dims[0]);
                    ACC v;
                    if (i0 < 0 || i1 < 0 || i2 < 0 || i3 < 0) v = (ACC)G.fill;
                    else v = load_elem<ACC>(img, dt, i0 + i1 * stride[1] + i2 * stride[2] + i3 * stride[3]);
                    if (op == B2F_WIN_MIN) acc = first ? v : (v < acc ? v : acc);
                    else if (op == B2F_WIN_MAX) acc = first ? v : (v > acc ? v : acc);
                    else acc = first ? v : add_rn<ACC>(acc, v);
                    first = false;
                }
            }
        }
    }
    if (op == B2F_WIN_MEAN) {
        if (odt == B2F_F32) ((float *)out)[o] = __fdiv_rn((float)acc, (float)G.wtotal);
        else ((double *)out)[o] = __ddiv_rn((double)acc, (double)G.wtotal);
        return;
    }
    store_elem<ACC>(out, odt, o, acc);
}

static long long numel(const b2f_array *a) {
    long long n = 1;
    for (int d = 0; d < a->ndim; ++d) n *= a->dims[d] < 0 ? 0 : a->dims[d];
    return n;
}

}  // namespace b2f

using namespace b2f;

extern "C" {

int b2f_findlocalextrema(const b2f_array *img, int32_t minima, const int64_t *window, const int32_t *edges,
                         int64_t *idx, int64_t cap, int64_t *count, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !window || !edges || !count || (cap > 0 && !idx)) return fail(B2F_EARG, "NULL argument");
    if (img->ndim < 1 || img->ndim > B2F_MAXDIM) return fail(B2F_ENOTSUP, "ndim %d not supported (1..%d)", img->ndim, B2F_MAXDIM);
    if (img->dtype < B2F_U8 || img->dtype > B2F_U32) return fail(B2F_EARG, "unsupported image dtype %d", img->dtype);
    PtGeom G;
    memset(&G, 0, sizeof G);
    G.ndim = img->ndim;
    G.n = numel(img);
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        G.dims[d] = d < img->ndim ? img->dims[d] : 1;
        if (d < img->ndim) {
            if (window[d] < 1) return fail(B2F_EARG, "window sizes must be positive");
            G.half[d] = (int)(window[d] >> 1);
            G.clip[d] = edges[d] ? 0 : 1;
        }
    }
    *count = 0;
    set_path("localextrema");
    if (G.n == 0) return 0;
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin;
    rc = stage_in(img, sin, st, true);
    if (rc) return rc;
    const long long nblocks = (G.n + PT_BLOCK - 1) / PT_BLOCK;
    if (nblocks > 0x7fffffffLL) { release(sin, st); return fail(B2F_ENOTSUP, "array too large"); }
    unsigned char *flags = nullptr;
    unsigned *counts = nullptr;
    long long *offsets = nullptr, *total = nullptr, *d_idx = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&flags, (size_t)G.n, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&counts, (size_t)nblocks * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&offsets, (size_t)(nblocks + 1) * sizeof(long long), st);
    total = offsets ? offsets + nblocks : nullptr;
    if (e == cudaSuccess && cap > 0) e = cudaMallocAsync((void **)&d_idx, (size_t)cap * sizeof(long long), st);
    if (e == cudaSuccess) {
        if (img->dtype == B2F_I64)
            pt_flag_kernel<long long><<<(unsigned)nblocks, PT_BLOCK, 0, st>>>(sin.dptr, img->dtype, G, minima, flags, counts);
        else
            pt_flag_kernel<double><<<(unsigned)nblocks, PT_BLOCK, 0, st>>>(sin.dptr, img->dtype, G, minima, flags, counts);
        pt_scan_kernel<<<1, 1024, 0, st>>>(counts, offsets, nblocks, total);
        count_launch(2);
        if (cap > 0) {
            pt_scatter_kernel<<<(unsigned)nblocks, PT_BLOCK, 0, st>>>(flags, offsets, G.n, d_idx, cap);
            count_launch();
        }
        e = cudaGetLastError();
    }
    long long h_total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_total, total, sizeof h_total, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && cap > 0 && h_total > 0) {
        const long long ncopy = h_total < cap ? h_total : cap;
        e = cudaMemcpyAsync(idx, d_idx, (size_t)ncopy * sizeof(long long), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (flags) cudaFreeAsync(flags, st);
    if (counts) cudaFreeAsync(counts, st);
    if (offsets) cudaFreeAsync(offsets, st);
    if (d_idx) cudaFreeAsync(d_idx, st);
    release(sin, st);
    if (e != cudaSuccess) return fail(B2F_ECUDA, "findlocalextrema failed: %s", cudaGetErrorString(e));
    *count = h_total;
    return 0;
}

int b2f_scale_into_slice(const b2f_array *src, const b2f_array *stack, int64_t slice, double scale, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!src || !stack) return fail(B2F_EARG, "NULL argument");
    if (stack->ndim != src->ndim + 1 || stack->ndim > B2F_MAXDIM) return fail(B2F_EDIM, "stack must have one leading axis more than src");
    for (int d = 0; d < src->ndim; ++d)
        if (stack->dims[d + 1] != src->dims[d]) return fail(B2F_EDIM, "stack axes do not match src axes");
    if (slice < 0 || slice >= stack->dims[0]) return fail(B2F_EDIM, "slice %lld outside the leading axis", (long long)slice);
    if (stack->dtype != B2F_F32 && stack->dtype != B2F_F64) return fail(B2F_EARG, "stack must be Float32 or Float64");
    if (src->mem != B2F_DEVICE || stack->mem != B2F_DEVICE) return fail(B2F_ENOTSUP, "scale_into_slice takes device arrays");
    const long long n = numel(src);
    set_path("scale_slice");
    if (n == 0) return 0;
    const long long blocks = (n + 255) / 256;
    pt_scale_slice_kernel<<<(unsigned)(blocks < sm_count() * 16 ? blocks : sm_count() * 16), 256, 0, st>>>(src->ptr, src->dtype, stack->ptr, stack->dtype, n,
                                                                                      stack->dims[0], slice, scale);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

int b2f_maxabs(const b2f_array *img, double *result, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !result) return fail(B2F_EARG, "NULL argument");
    const long long n = numel(img);
    if (n == 0) return fail(B2F_EARG, "reducing over an empty collection is not allowed");
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin;
    rc = stage_in(img, sin, st, true);
    if (rc) return rc;
    unsigned long long *acc = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&acc, 16, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(acc, 0, 16, st);
    unsigned long long h[2] = {0, 0};
    if (e == cudaSuccess) {
        const long long blocks = (n + 255) / 256;
        pt_maxabs_kernel<<<(unsigned)(blocks < sm_count() * 16 ? blocks : sm_count() * 16), 256, 0, st>>>(sin.dptr, img->dtype, n, acc, (int *)(acc + 1));
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (acc) cudaFreeAsync(acc, st);
    release(sin, st);
    if (e != cudaSuccess) return fail(B2F_ECUDA, "maxabs failed: %s", cudaGetErrorString(e));
    double m;
    memcpy(&m, &h[0], sizeof m);
    *result = (h[1] & 1ULL) ? __builtin_nan("") : m;
    return 0;
}

int b2f_gather(const b2f_array *arr, const int64_t *idx, int64_t n, double *values, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!arr || (n > 0 && (!idx || !values))) return fail(B2F_EARG, "NULL argument");
    if (n <= 0) return 0;
    const long long total = numel(arr);
    for (int64_t i = 0; i < n; ++i)
        if (idx[i] < 0 || idx[i] >= total) return fail(B2F_EDIM, "index %lld outside the array", (long long)idx[i]);
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin;
    rc = stage_in(arr, sin, st, true);
    if (rc) return rc;
    long long *d_idx = nullptr;
    double *d_out = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_idx, (size_t)n * 8, st);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&d_out, (size_t)n * 8, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, idx, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        pt_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sin.dptr, arr->dtype, d_idx, n, d_out);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, d_out, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (d_idx) cudaFreeAsync(d_idx, st);
    if (d_out) cudaFreeAsync(d_out, st);
    release(sin, st);
    if (e != cudaSuccess) return fail(B2F_ECUDA, "gather failed: %s", cudaGetErrorString(e));
    return 0;
}

int b2f_na_prepare(const b2f_array *img, int32_t na_mode, const b2f_array *imgtmp, const b2f_array *valid, int32_t *hasna,
                   void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !hasna) return fail(B2F_EARG, "NULL argument");
    if (na_mode < 0 || na_mode > 2) return fail(B2F_EARG, "na_mode must be 0 (isnan), 1 (!isfinite) or 2 (never)");
    const long long n = numel(img);
    for (const b2f_array *a : {imgtmp, valid}) {
        if (!a) continue;
        if (a->dtype != B2F_F32 && a->dtype != B2F_F64) return fail(B2F_EARG, "imgtmp / valid must be Float32 or Float64");
        if (numel(a) != n) return fail(B2F_EDIM, "imgtmp / valid must have the axes of img");
        if (a->mem != img->mem) return fail(B2F_ENOTSUP, "img, imgtmp and valid must live in the same memory space");
    }
    *hasna = 0;
    set_path("na_prepare");
    if (n == 0) return 0;
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin, st1, st2;
    rc = stage_in(img, sin, st, true);
    if (!rc && imgtmp) rc = stage_in(imgtmp, st1, st, false);
    if (!rc && valid) rc = stage_in(valid, st2, st, false);
    int *flag = nullptr;
    cudaError_t e = rc ? cudaSuccess : cudaMallocAsync((void **)&flag, 4, st);
    if (!rc && e == cudaSuccess) e = cudaMemsetAsync(flag, 0, 4, st);
    if (!rc && e == cudaSuccess) {
        const long long blocks = (n + 255) / 256;
        pt_na_prepare_kernel<<<(unsigned)(blocks < sm_count() * 16 ? blocks : sm_count() * 16), 256, 0, st>>>(
            sin.dptr, img->dtype, n, na_mode, imgtmp ? st1.dptr : nullptr, imgtmp ? imgtmp->dtype : 0, valid ? st2.dptr : nullptr,
            valid ? valid->dtype : 0, flag);
        count_launch();
        e = cudaGetLastError();
    }
    int h = 0;
    if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, st);
    if (!rc && e == cudaSuccess && imgtmp && imgtmp->mem == B2F_HOST) e = cudaMemcpyAsync(imgtmp->ptr, st1.dptr, st1.bytes, cudaMemcpyDeviceToHost, st);
    if (!rc && e == cudaSuccess && valid && valid->mem == B2F_HOST) e = cudaMemcpyAsync(valid->ptr, st2.dptr, st2.bytes, cudaMemcpyDeviceToHost, st);
    if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (flag) cudaFreeAsync(flag, st);
    release(sin, st); release(st1, st); release(st2, st);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(B2F_ECUDA, "na_prepare failed: %s", cudaGetErrorString(e));
    *hasna = h;
    return 0;
}

int b2f_divide(const b2f_array *out, const b2f_array *den, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out || !den) return fail(B2F_EARG, "NULL argument");
    if ((out->dtype != B2F_F32 && out->dtype != B2F_F64) || (den->dtype != B2F_F32 && den->dtype != B2F_F64))
        return fail(B2F_ENOTSUP, "divide takes Float32 / Float64 arrays");
    const long long n = numel(out);
    if (numel(den) != n) return fail(B2F_EDIM, "out and den must have the same axes");
    set_path("divide");
    if (n == 0) return 0;
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged so, sd;
    rc = stage_in(out, so, st, true);
    if (!rc) rc = stage_in(den, sd, st, true);
    cudaError_t e = cudaSuccess;
    if (!rc) {
        const long long blocks = (n + 255) / 256;
        pt_divide_kernel<<<(unsigned)(blocks < sm_count() * 16 ? blocks : sm_count() * 16), 256, 0, st>>>(so.dptr, out->dtype, sd.dptr, den->dtype, n);
        count_launch();
        e = cudaGetLastError();
        if (e == cudaSuccess && out->mem == B2F_HOST) {
            e = cudaMemcpyAsync(out->ptr, so.dptr, so.bytes, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
    }
    release(so, st); release(sd, st);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(B2F_ECUDA, "divide failed: %s", cudaGetErrorString(e));
    return 0;
}

int b2f_normalize_dims(const b2f_array *out, const double *const *factors, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!out || !factors) return fail(B2F_EARG, "NULL argument");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "normalize_dims takes a Float32 / Float64 array");
    if (out->ndim < 1 || out->ndim > B2F_MAXDIM) return fail(B2F_ENOTSUP, "ndim %d not supported", out->ndim);
    const long long n = numel(out);
    set_path("normalize_dims");
    if (n == 0) return 0;
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged so;
    rc = stage_in(out, so, st, true);
    if (rc) return rc;
    PtFactors F;
    memset(&F, 0, sizeof F);
    F.ndim = out->ndim;
    long long total = 0;
    for (int d = 0; d < out->ndim; ++d) { F.dims[d] = out->dims[d]; total += out->dims[d]; if (!factors[d]) { release(so, st); return fail(B2F_EARG, "NULL factor"); } }
    double *d_f = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d_f, (size_t)total * 8, st);
    long long off = 0;
    for (int d = 0; d < out->ndim && e == cudaSuccess; ++d) {
        e = cudaMemcpyAsync(d_f + off, factors[d], (size_t)out->dims[d] * 8, cudaMemcpyHostToDevice, st);
        F.f[d] = d_f + off;
        off += out->dims[d];
    }
    if (e == cudaSuccess) {
        const long long blocks = (n + 255) / 256;
        pt_normalize_dims_kernel<<<(unsigned)(blocks < sm_count() * 16 ? blocks : sm_count() * 16), 256, 0, st>>>(so.dptr, out->dtype, n, F);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && out->mem == B2F_HOST) e = cudaMemcpyAsync(out->ptr, so.dptr, so.bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);       // the factor vectors are host memory of the caller
    if (d_f) cudaFreeAsync(d_f, st);
    release(so, st);
    if (e != cudaSuccess) return fail(B2F_ECUDA, "normalize_dims failed: %s", cudaGetErrorString(e));
    return 0;
}

// shared by b2f_mapwindow_median and b2f_mapwindow_reduce
static int mapwindow_reduce_impl(const b2f_array *img, const b2f_array *out, int op, const int64_t *win_lo, const int64_t *win_hi,
                                 const b2f_border *border, const int64_t *idx_first, const int64_t *idx_step, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !out || !win_lo || !win_hi || !border) return fail(B2F_EARG, "NULL argument");
    if ((idx_first == nullptr) != (idx_step == nullptr)) return fail(B2F_EARG, "idx_first and idx_step go together");
    if (op < B2F_WIN_MEDIAN || op > B2F_WIN_MAX) return fail(B2F_EARG, "unknown window reduction %d", op);
    const int N = img->ndim;
    if (N < 1 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "mapwindow needs 1..%d dims and equal rank", B2F_MAXDIM);
    if (img->dtype == B2F_N0F8) return fail(B2F_ENOTSUP, "window reductions of N0f8 images are not available");
    const bool isf = img->dtype == B2F_F32 || img->dtype == B2F_F64;
    int want;
    switch (op) {
        case B2F_WIN_MEDIAN: case B2F_WIN_MEAN: want = img->dtype == B2F_F32 ? B2F_F32 : B2F_F64; break;
        case B2F_WIN_SUM: want = isf ? img->dtype : B2F_I64; break;
        default: want = img->dtype; break;
    }
    if (out->dtype != want) return fail(B2F_EARG, "output eltype %d does not match the reduction's result type %d", out->dtype, want);
    if (border->style > B2F_INNER) return fail(B2F_ENOTSUP, "border style %d is not supported by mapwindow", border->style);
    PtWin G;
    memset(&G, 0, sizeof G);
    G.ndim = N;
    G.style = border->style == B2F_INNER ? B2F_REPLICATE : border->style;
    G.fill = border->fill;
    G.nout = 1;
    G.wtotal = 1;
    Box ia = axes_of(img), oa = axes_of(out);
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        G.dims[d] = d < N ? img->dims[d] : 1;
        G.odims[d] = d < N ? out->dims[d] : 1;
        G.ostep[d] = (d < N && idx_step) ? idx_step[d] : 1;
        G.ooff[d] = d < N ? ((idx_first ? idx_first[d] : oa.lo[d]) - ia.lo[d]) : 0;
        G.wlo[d] = d < N ? (int)win_lo[d] : 0;
        G.wn[d] = d < N ? (int)(win_hi[d] - win_lo[d] + 1) : 1;
        if (G.wn[d] < 1) return fail(B2F_EARG, "empty window");
        if (G.ostep[d] < 1) return fail(B2F_EARG, "indices must be increasing ranges");
        G.wtotal *= G.wn[d];
        G.nout *= G.odims[d] < 0 ? 0 : G.odims[d];
        if (d < N && G.odims[d] > 0) {
            const long long first = G.ooff[d], last = G.ooff[d] + (G.odims[d] - 1) * G.ostep[d];
            if (first < 0 || last > G.dims[d] - 1) return fail(B2F_EDIM, "requested indices exceed the image axes");
            if (border->style == B2F_INNER && (first + win_lo[d] < 0 || last + win_hi[d] > G.dims[d] - 1))
                return fail(B2F_EDIM, "requested indices are not in the interior for Inner()");
            if (border->style != B2F_FILL && (win_lo[d] > 0 || win_hi[d] < 0) && border->style != B2F_INNER)
                return fail(B2F_ENOTSUP, "windows that do not contain their centre need Fill or Inner borders here");
        }
    }
    if (op == B2F_WIN_MEDIAN && G.wtotal > PT_MAXWIN) return fail(B2F_ENOTSUP, "median windows hold at most %d elements", PT_MAXWIN);
    set_path(op == B2F_WIN_MEDIAN ? "median" : "winreduce");
    if (G.nout == 0 || numel(img) == 0) return 0;
    int rc = ensure_ctx();
    if (rc) return rc;
    Staged sin, so;
    rc = stage_in(img, sin, st, true);
    if (!rc) rc = stage_in(out, so, st, false);
    cudaError_t e = cudaSuccess;
    if (!rc) {
        const long long blocks = (G.nout + 127) / 128;
        if (op == B2F_WIN_MEDIAN) {
            if (img->dtype == B2F_I64) pt_median_kernel<long long><<<(unsigned)blocks, 128, 0, st>>>(sin.dptr, img->dtype, so.dptr, out->dtype, G);
            else pt_median_kernel<double><<<(unsigned)blocks, 128, 0, st>>>(sin.dptr, img->dtype, so.dptr, out->dtype, G);
        } else if (img->dtype == B2F_F32) {
            pt_winreduce_kernel<float><<<(unsigned)blocks, 128, 0, st>>>(sin.dptr, img->dtype, so.dptr, out->dtype, op, G);
        } else if (img->dtype == B2F_F64) {
            pt_winreduce_kernel<double><<<(unsigned)blocks, 128, 0, st>>>(sin.dptr, img->dtype, so.dptr, out->dtype, op, G);
        } else {
            pt_winreduce_kernel<long long><<<(unsigned)blocks, 128, 0, st>>>(sin.dptr, img->dtype, so.dptr, out->dtype, op, G);
        }
        count_launch();
        e = cudaGetLastError();
        if (e == cudaSuccess && out->mem == B2F_HOST) {
            e = cudaMemcpyAsync(out->ptr, so.dptr, so.bytes, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        } else if (e == cudaSuccess && img->mem == B2F_HOST) {
            e = cudaStreamSynchronize(st);
        }
    }
    release(sin, st); release(so, st);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(B2F_ECUDA, "mapwindow reduction failed: %s", cudaGetErrorString(e));
    return 0;
}

int b2f_mapwindow_median(const b2f_array *img, const b2f_array *out, const int64_t *win_lo, const int64_t *win_hi,
                         const b2f_border *border, void *stream) {
    return mapwindow_reduce_impl(img, out, B2F_WIN_MEDIAN, win_lo, win_hi, border, nullptr, nullptr, stream);
}

int b2f_mapwindow_reduce(const b2f_array *img, const b2f_array *out, int32_t op, const int64_t *win_lo, const int64_t *win_hi,
                         const b2f_border *border, const int64_t *idx_first, const int64_t *idx_step, void *stream) {
    return mapwindow_reduce_impl(img, out, op, win_lo, win_hi, border, idx_first, idx_step, stream);
}

}  // extern "C"
