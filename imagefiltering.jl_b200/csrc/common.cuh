// common.cuh — shared host/device helpers of libb2f.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/b2f.h"

namespace b2f {

// ---- host-side status plumbing ---------------------------------------------------------------
int fail(int code, const char *fmt, ...);
// multiprocessor count of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200)
int sm_count();
void set_path(const char *name);
void count_launch(int n = 1);
int accum_mode();            // B2F_ACCUM_* of the calling thread (b2f_set_accum_mode)
#define B2F_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ::b2f::fail(B2F_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));    \
    } while (0)

// stream-ordered frees on scope exit: an early error return (B2F_CUDA) releases what was allocated so far
struct AsyncFrees {
    cudaStream_t st;
    std::vector<void *> v;
    explicit AsyncFrees(cudaStream_t s) : st(s) {}
    AsyncFrees(const AsyncFrees &) = delete;
    AsyncFrees &operator=(const AsyncFrees &) = delete;
    void push_back(void *p) { v.push_back(p); }
    ~AsyncFrees() {
        for (void *p : v)
            if (p) cudaFreeAsync(p, st);
    }
};

// ---- geometry ------------------------------------------------------------------------------------
struct Box {  // inclusive index ranges; axes >= ndim are 0:0
    int64_t lo[B2F_MAXDIM], hi[B2F_MAXDIM];
    __host__ __device__ int64_t len(int d) const { return hi[d] - lo[d] + 1; }
    bool empty() const {
        for (int d = 0; d < B2F_MAXDIM; ++d)
            if (hi[d] < lo[d]) return true;
        return false;
    }
    int64_t count() const {
        int64_t n = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) n *= (hi[d] < lo[d] ? 0 : len(d));
        return n;
    }
};

inline Box axes_of(const b2f_array *a) {
    Box b;
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        if (d < a->ndim) {
            b.lo[d] = a->origin[d];
            b.hi[d] = a->origin[d] + a->dims[d] - 1;
        } else {
            b.lo[d] = b.hi[d] = 0;
        }
    }
    return b;
}

inline size_t dtype_size(int dt) {
    switch (dt) {
        case B2F_U8: case B2F_N0F8: return 1;
        case B2F_I16: case B2F_U16: return 2;
        case B2F_I32: case B2F_U32: case B2F_F32: return 4;
        default: return 8;
    }
}
inline bool is_int_dtype(int dt) {
    return dt == B2F_U8 || dt == B2F_I16 || dt == B2F_I32 || dt == B2F_I64 || dt == B2F_U16 || dt == B2F_U32;
}
inline bool int_range(int dt, int64_t &lo, int64_t &hi) {
    switch (dt) {
        case B2F_U8: lo = 0; hi = 255; return true;
        case B2F_I16: lo = -32768; hi = 32767; return true;
        case B2F_U16: lo = 0; hi = 65535; return true;
        case B2F_I32: lo = INT32_MIN; hi = INT32_MAX; return true;
        case B2F_U32: lo = 0; hi = UINT32_MAX; return true;
        case B2F_I64: lo = INT64_MIN; hi = INT64_MAX; return true;
    }
    return false;
}

// A stage after validation: tap index ranges per axis and the copy-kernel test.
struct StageInfo {
    const b2f_stage *s;
    int64_t lo[B2F_MAXDIM], hi[B2F_MAXDIM];
    bool copy;
};

// The resolved call: what the reference's steps 5-6 (src/imfilter.jl:321-341, src/border.jl:614-684)
// would have produced, without materialising anything.
struct Plan {
    int ndim;
    std::vector<StageInfo> stages;
    std::vector<int> active;                 // non-copy stages, in order
    int64_t pad_lo[B2F_MAXDIM], pad_hi[B2F_MAXDIM];
    Box img_ax, out_ax, roi, padded_ax;
    std::vector<Box> region;                 // region[a] = output region of active stage a
    int style;
    double fill;                             // already converted through eltype(img)
};

int make_plan(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int nstages,
              const b2f_border *border, const int64_t *roi_lo, const int64_t *roi_hi, Plan &P);

// ---- device-side element access --------------------------------------------------------------------
// N0f8 -> double, correctly rounded i/255 without a division: q = i*r, one Newton correction with two
// FMAs (Markstein).  tests/test_host_logic.py checks all 256 values against IEEE division.
__host__ __device__ inline double n0f8_to_f64(unsigned i) {
    const double r = 1.0 / 255.0;
    const double x = (double)i;
#ifdef __CUDA_ARCH__
    double q = x * r;
    double rem = fma(-q, 255.0, x);
    return fma(rem, r, q);
#else
    double q = x * r;
    double rem = __builtin_fma(-q, 255.0, x);
    return __builtin_fma(rem, r, q);
#endif
}
__host__ __device__ inline float n0f8_to_f32(unsigned i) {
    const float r = 1.0f / 255.0f;
    const float x = (float)i;
#ifdef __CUDA_ARCH__
    float q = x * r;
    float rem = fmaf(-q, 255.0f, x);
    return fmaf(rem, r, q);
#else
    float q = x * r;
    float rem = __builtin_fmaf(-q, 255.0f, x);
    return __builtin_fmaf(rem, r, q);
#endif
}

template <typename CT> struct N0f8Conv;
template <> struct N0f8Conv<double> { __device__ static double f(unsigned i) { return n0f8_to_f64(i); } };
template <> struct N0f8Conv<float> { __device__ static float f(unsigned i) { return n0f8_to_f32(i); } };
template <> struct N0f8Conv<long long> { __device__ static long long f(unsigned i) { return (long long)i; } };

// load element `i` of an array of runtime dtype, converted to the compute type
template <typename CT>
__device__ __forceinline__ CT load_elem(const void *__restrict__ p, int dt, int64_t i) {
    switch (dt) {
        case B2F_U8: return (CT)((const uint8_t *)p)[i];
        case B2F_N0F8: return N0f8Conv<CT>::f(((const uint8_t *)p)[i]);
        case B2F_I16: return (CT)((const int16_t *)p)[i];
        case B2F_U16: return (CT)((const uint16_t *)p)[i];
        case B2F_I32: return (CT)((const int32_t *)p)[i];
        case B2F_U32: return (CT)((const uint32_t *)p)[i];
        case B2F_I64: return (CT)((const long long *)p)[i];
        case B2F_F32: return (CT)((const float *)p)[i];
        default: return (CT)((const double *)p)[i];
    }
}

// store with conversion; returns false when the value is not representable (InexactError)
template <typename CT>
__device__ __forceinline__ bool store_elem(void *__restrict__ p, int dt, int64_t i, CT v) {
    switch (dt) {
        case B2F_F64: ((double *)p)[i] = (double)v; return true;
        case B2F_F32: ((float *)p)[i] = (float)v; return true;
        default: break;
    }
    long long iv = (long long)v;
    if ((CT)iv != v) return false;
    switch (dt) {
        case B2F_U8: if (iv < 0 || iv > 255) return false; ((uint8_t *)p)[i] = (uint8_t)iv; return true;
        case B2F_I16: if (iv < -32768 || iv > 32767) return false; ((int16_t *)p)[i] = (int16_t)iv; return true;
        case B2F_U16: if (iv < 0 || iv > 65535) return false; ((uint16_t *)p)[i] = (uint16_t)iv; return true;
        case B2F_I32: if (iv < INT32_MIN || iv > INT32_MAX) return false; ((int32_t *)p)[i] = (int32_t)iv; return true;
        case B2F_U32: if (iv < 0 || iv > (long long)UINT32_MAX) return false; ((uint32_t *)p)[i] = (uint32_t)iv; return true;
        case B2F_I64: ((long long *)p)[i] = iv; return true;
    }
    return false;
}

// Border index remap (reference src/border.jl:564-590 padindex, :644-645 modrange), 0-based position
// `i` relative to an axis of length n.  Returns -1 for B2F_FILL outside the array.  Folds fully (pads
// larger than the array), like padarray.
__host__ __device__ __forceinline__ int64_t remap_index(int style, int64_t i, int64_t n) {
    if (i >= 0 && i < n) return i;
    switch (style) {
        case B2F_REPLICATE: return i < 0 ? 0 : n - 1;
        case B2F_CIRCULAR: { int64_t m = i % n; return m < 0 ? m + n : m; }
        case B2F_SYMMETRIC: {
            const int64_t p = 2 * n;
            int64_t m = i % p; if (m < 0) m += p;
            return m < n ? m : p - 1 - m;
        }
        case B2F_REFLECT: {
            const int64_t p = 2 * n - 2;
            int64_t m = i % p; if (m < 0) m += p;
            return m < n ? m : p - m;
        }
        default: return -1;
    }
}

// the multiply-accumulate the reference performs: separate multiply and add for exact (double / int)
// modes, FMA for float
template <typename CT> __device__ __forceinline__ CT mac(CT acc, CT a, CT k);
template <> __device__ __forceinline__ double mac<double>(double acc, double a, double k) {
    return __dadd_rn(acc, __dmul_rn(a, k));
}
template <> __device__ __forceinline__ float mac<float>(float acc, float a, float k) { return fmaf(a, k, acc); }
template <> __device__ __forceinline__ long long mac<long long>(long long acc, long long a, long long k) {
    return acc + a * k;
}

// never-contracted add / multiply (the Laplacian loop of src/specialty.jl:3-16 is mul, then adds)
template <typename CT> __device__ __forceinline__ CT add_rn(CT a, CT b) { return a + b; }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <typename CT> __device__ __forceinline__ CT mul_rn(CT a, CT b) { return a * b; }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }

// ---- host staging shared by the ABI files (api.cu, points.cu) -------------------------------------------
struct Staged {  // device view of one array of the call
    void *dptr = nullptr;
    bool owned = false;
    size_t bytes = 0;
};
int ensure_ctx();
int stage_in(const b2f_array *a, Staged &s, cudaStream_t st, bool copy);   // host arrays: pool allocation (+ H2D copy)
void release(Staged &s, cudaStream_t st);

// ---- kernel-family entry points (host) ---------------------------------------------------------------
struct DevArrays {        // device-resident views of the call's arrays
    const void *img; int img_dt;
    void *out; int out_dt;
};

// generic per-stage path: any ndim <= 4, any dtype, any cascade
int run_generic(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st);
// fused 2-D separable (axes 0 and 1, two 1-D stages), up to 2 planes sharing one input
bool fused2d_applicable(const Plan *plans, int nplanes, int img_dt, const int *out_dt);
int run_fused2d(const Plan *plans, int nplanes, const void *d_img, int img_dt, void *const *d_outs,
                const int *out_dt, cudaStream_t st);
// warp-streamed fused 2-D separable (the fast path for <= 16 taps, x-then-y order)
bool stream2d_applicable(const Plan *plans, int nplanes, int img_dt, const int *out_dt);
int run_stream2d(const Plan *plans, int nplanes, const void *d_img, int img_dt, void *const *d_outs,
                 const int *out_dt, cudaStream_t st);
// N-d separable cascade as a chain of streamed passes (also the 3-D path)
bool sepnd_applicable(const Plan &P, int img_dt, int out_dt);
int run_sepnd(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st);
// fused 3-D separable cascade (x, y, z in one launch; Float32), also in slab form (planes in own/lo/hi buffers)
bool stream3d_applicable(const Plan &P, int img_dt, int out_dt);
bool stream3d_xy_capable(const Plan &P);          // ... and every slab of it can take xy-filtered boundary planes (TMA path)
int run_stream3d(const Plan &P, const void *d_img, void *d_out, cudaStream_t st);
int run_stream3d_slab(const Plan &P, const void *own, const void *lo, int64_t lo_n, const void *hi, int64_t hi_n,
                      int64_t own_first, int64_t own_n, void *d_out, cudaStream_t st, const void *flag_lo = nullptr,
                      const void *flag_hi = nullptr, int epoch = 0, int lo_early_rows = 0, const b2f_slab_xy *xy = nullptr);
// dense 2-D single stage (register-blocked FMA)
bool dense2d_applicable(const Plan &P, int img_dt, int out_dt);
int run_dense2d(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st);
// dense 3-D single stage (shared-memory block, 4 outputs per thread)
bool dense3d_applicable(const Plan &P, int img_dt, int out_dt);
int run_dense3d(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st);
// running extrema
int run_extrema(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                cudaStream_t st);

}  // namespace b2f
