// extrema_vh.cu — K4 for ANY window width and element type: separable running minimum / maximum by the van Herk /
// Gil-Werman block recurrence, O(1) comparisons per element and axis whatever the window.
//
// Replaces, for windows wider than the register kernel (extrema2d.cuh: Float32, <= 16) and for every other eltype
// (UInt8 / N0f8 images, Int16 ... Float64), the reference's Lemire wedge per axis + permutedims (src/mapwindow.jl:388-473)
// and the O(prod(w)) window copy of `minimum` / `maximum` (src/mapwindow.jl:270-333).  The round-1 fallback for these
// cases was an O(prod(w)) scan per pixel (generic.cu).
//
// One launch = one axis pass of one reduction (min or max) over a dense array viewed as [P | N | Q] (P = product of the
// faster axes, N = the filtered axis, Q = the slower ones).  A CTA stages a tile of VH_C = 32 columns x S_in positions in
// shared memory as X[s][c]; out[i] = op(x[i+lo .. i+lo+w-1]) (window truncated at the array ends == positions outside the
// array hold the identity of op; B2F_FILL: they hold the fill value, src/mapwindow.jl:326-333):
//   phase 1   every (column, block of w positions): suffix scan  H[s] = op(x[s], H[s+1])           (backward, in H)
//   phase 2   every (column, output block b): for t = 0..w-1:  out[bw+t] = op(H[bw+t], run);  run = op(run, x[(b+1)w+t])
//             (run = prefix of the NEXT block, built on the fly) — written back into H, so that
//   phase 3   the tile leaves shared memory with coalesced stores.
// Comparisons are the reference's strict `<` / `>` on the element type (integers exact; floats bit-identical for NaN-free
// data).  extrema = a min pass and a max pass per axis; the last pass of each writes its half of the (min,max) tuples.
#include <algorithm>
#include <limits>

#include "common.cuh"

namespace b2f {

constexpr int VH_C = 32;          // columns per tile
constexpr int VH_CP = VH_C + 1;   // shared-memory pitch (odd: conflict-free along either index)
constexpr int VH_NT = 256;

template <typename T>
struct VHParams {
    const T *src;
    T *dst;
    long long P, N, Q;            // source / destination view (same shape)
    long long dst_es, dst_off;    // destination element stride / offset in elements (2 / 0|1: halves of Tuple{T,T})
    int lo, w;                    // window = [i + lo, i + lo + w - 1]
    int fill_on;
    T fill;
    int S, S_in;                  // outputs per tile (multiple of w), staged positions per tile (S + w - 1 rounded up to blocks)
    long long nseg, ngrp, pgroups;   // tiles along N; column groups; column groups per q (P > 1)
};

template <typename T, bool MAXMODE> struct VHOp {
    __device__ static __forceinline__ T ident() { return MAXMODE ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max(); }
    __device__ static __forceinline__ T op(T a, T b) { return MAXMODE ? (b > a ? b : a) : (b < a ? b : a); }
};
template <bool MAXMODE> struct VHOp<float, MAXMODE> {
    __device__ static __forceinline__ float ident() { return MAXMODE ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000); }
    __device__ static __forceinline__ float op(float a, float b) { return MAXMODE ? (b > a ? b : a) : (b < a ? b : a); }
};
template <bool MAXMODE> struct VHOp<double, MAXMODE> {
    __device__ static __forceinline__ double ident() {
        return MAXMODE ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL);
    }
    __device__ static __forceinline__ double op(double a, double b) { return MAXMODE ? (b > a ? b : a) : (b < a ? b : a); }
};

template <typename T, bool MAXMODE, bool AXIS0>
__global__ void __launch_bounds__(VH_NT) vh_pass_kernel(const VHParams<T> p) {
    extern __shared__ __align__(16) unsigned char vh_smem[];
    T *X = reinterpret_cast<T *>(vh_smem);               // [S_in][VH_CP]
    T *H = X + (size_t)p.S_in * VH_CP;
    typedef VHOp<T, MAXMODE> Op;
    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const long long seg = tile % p.nseg, grp = tile / p.nseg;
    const long long i0 = seg * p.S;                       // first output position of the tile
    const long long s0 = i0 + p.lo;                       // global position of staged index 0
    // columns of this tile: AXIS0 (P == 1): 32 lines q0 .. q0+31; else 32 consecutive p of one q
    long long q0, pbase;
    int ncol;
    if (AXIS0) {
        q0 = grp * VH_C; pbase = 0;
        ncol = (int)min((long long)VH_C, p.Q - q0);
    } else {
        q0 = grp / p.pgroups; pbase = (grp % p.pgroups) * VH_C;
        ncol = (int)min((long long)VH_C, p.P - pbase);
    }
    const T outside = p.fill_on ? p.fill : Op::ident();
    const int n_in = p.S_in;
    // ---- stage: X[s][c] = x(s0 + s) or the identity / fill outside the array -------------------------------------------
    if (AXIS0) {
        for (int e = tid; e < n_in * VH_C; e += VH_NT) {
            const int c = e / n_in, s = e - c * n_in;
            T v = outside;
            const long long g = s0 + s;
            if (c < ncol && g >= 0 && g < p.N) v = p.src[(q0 + c) * p.N + g];
            X[s * VH_CP + c] = v;
        }
    } else {
        for (int e = tid; e < n_in * VH_C; e += VH_NT) {
            const int s = e / VH_C, c = e - s * VH_C;
            T v = outside;
            const long long g = s0 + s;
            if (c < ncol && g >= 0 && g < p.N) v = p.src[(q0 * p.N + g) * p.P + pbase + c];
            X[s * VH_CP + c] = v;
        }
    }
    __syncthreads();
    const int w = p.w, nbi = n_in / w;                   // S_in is a multiple of w
    // ---- phase 1: suffix scans inside every block of w staged positions --------------------------------------------------
    for (int task = tid; task < nbi * VH_C; task += VH_NT) {
        const int c = task % VH_C, b = task / VH_C;
        T run = X[((b + 1) * w - 1) * VH_CP + c];
        H[((b + 1) * w - 1) * VH_CP + c] = run;
        for (int s = (b + 1) * w - 2; s >= b * w; --s) {
            run = Op::op(run, X[s * VH_CP + c]);
            H[s * VH_CP + c] = run;
        }
    }
    __syncthreads();
    // ---- phase 2: out[bw + t] = op(suffix of block b from t, prefix of block b+1 up to t-1); result kept in H -----------
    const int nbo = p.S / w;
    for (int task = tid; task < nbo * VH_C; task += VH_NT) {
        const int c = task % VH_C, b = task / VH_C;
        T run = Op::ident();
        bool have = false;
        for (int t = 0; t < w; ++t) {
            const int s = b * w + t;
            const T h = H[s * VH_CP + c];
            H[s * VH_CP + c] = have ? Op::op(h, run) : h;
            const T x = X[(s + w) * VH_CP + c];
            run = have ? Op::op(run, x) : x;
            have = true;
        }
    }
    __syncthreads();
    // ---- phase 3: coalesced stores of the S outputs of every column ------------------------------------------------------
    if (AXIS0) {
        for (int e = tid; e < p.S * VH_C; e += VH_NT) {
            const int c = e / p.S, s = e - c * p.S;
            const long long g = i0 + s;
            if (c < ncol && g < p.N) p.dst[((q0 + c) * p.N + g) * p.dst_es + p.dst_off] = H[s * VH_CP + c];
        }
    } else {
        for (int e = tid; e < p.S * VH_C; e += VH_NT) {
            const int s = e / VH_C, c = e - s * VH_C;
            const long long g = i0 + s;
            if (c < ncol && g < p.N) p.dst[((q0 * p.N + g) * p.P + pbase + c) * p.dst_es + p.dst_off] = H[s * VH_CP + c];
        }
    }
}

// crop of a dense array to a box, into a destination with an element stride / offset (the last step for Inner() outputs)
template <typename T>
__global__ void __launch_bounds__(256) vh_crop_kernel(const T *src, T *dst, long long es, long long off, long long s0, long long s1,
                                                      long long s2, long long o0, long long o1, long long o2, long long o3, long long f0,
                                                      long long f1, long long f2, long long f3) {
    const long long total = o0 * o1 * o2 * o3;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        long long r = t;
        const long long a = r % o0; r /= o0;
        const long long b = r % o1; r /= o1;
        const long long c = r % o2; r /= o2;
        dst[t * es + off] = src[(a + f0) + s0 * ((b + f1) + s1 * ((c + f2) + s2 * (r + f3)))];
    }
}

template <typename T>
static int vh_pass(const T *src, T *dst, long long es, long long off, const int64_t *dims, int ndim, int axis, int lo, int w, bool maxmode,
                   bool fill_on, double fill, cudaStream_t st) {
    VHParams<T> p;
    memset(&p, 0, sizeof p);
    p.src = src; p.dst = dst; p.dst_es = es; p.dst_off = off;
    p.P = 1; p.Q = 1;
    for (int d = 0; d < axis; ++d) p.P *= dims[d];
    for (int d = axis + 1; d < ndim; ++d) p.Q *= dims[d];
    p.N = dims[axis];
    p.lo = lo; p.w = w; p.fill_on = fill_on ? 1 : 0; p.fill = (T)fill;
    // tile: as many staged positions as 2 arrays of S_in x 33 elements fit in `budget` bytes; at least two blocks
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    size_t budget = 96 * 1024;
    auto fit = [&](size_t bytes) { return (long long)(bytes / (2 * VH_CP * sizeof(T))); };
    long long cap = fit(budget);
    if (cap < 2LL * w) { budget = (size_t)max_optin; cap = fit(budget); }
    if (cap < 2LL * w) return fail(B2F_ENOTSUP, "window of %d elements exceeds the shared-memory tile of the running-extrema kernel", w);
    long long nb = cap / w;                                  // staged blocks
    const long long need = (p.N + w - 1) / w + 1;            // blocks that cover the whole axis in one tile
    nb = std::min(nb, std::max(2LL, need));
    p.S_in = (int)(nb * w);
    p.S = (int)((nb - 1) * w);
    p.nseg = (p.N + p.S - 1) / p.S;
    const bool axis0 = p.P == 1;
    if (axis0) { p.pgroups = 1; p.ngrp = (p.Q + VH_C - 1) / VH_C; }
    else { p.pgroups = (p.P + VH_C - 1) / VH_C; p.ngrp = p.pgroups * p.Q; }
    const long long blocks = p.nseg * p.ngrp;
    if (blocks > 0x7fffffffLL) return fail(B2F_ENOTSUP, "running-extrema grid too large");
    const size_t smem = 2 * (size_t)p.S_in * VH_CP * sizeof(T);
#define VH_LAUNCH(MAXM, AX0)                                                                                          \
    do {                                                                                                              \
        auto kern = vh_pass_kernel<T, MAXM, AX0>;                                                                     \
        B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(smem, (size_t)48 * 1024))); \
        kern<<<(unsigned)blocks, VH_NT, smem, st>>>(p);                                                              \
    } while (0)
    if (maxmode) { if (axis0) VH_LAUNCH(true, true); else VH_LAUNCH(true, false); }
    else { if (axis0) VH_LAUNCH(false, true); else VH_LAUNCH(false, false); }
#undef VH_LAUNCH
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int run_extrema_vh_typed(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved, const Box &out_ax,
                                const int64_t *wlo, const int64_t *whi, int style, double fill, cudaStream_t st) {
    const int nd = img->ndim;
    Box ia = axes_of(img);
    int64_t dims[B2F_MAXDIM];
    int64_t total = 1;
    bool full = true;
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        dims[d] = d < nd ? ia.len(d) : 1;
        total *= dims[d];
        if (d < nd && (out_ax.lo[d] != ia.lo[d] || out_ax.hi[d] != ia.hi[d])) full = false;
    }
    std::vector<int> axes;
    for (int d = 0; d < nd; ++d)
        if (whi[d] - wlo[d] + 1 > 1 || wlo[d] != 0) axes.push_back(d);
    AsyncFrees tmp(st);
    auto temp = [&](T *&ptr) -> int {
        void *q = nullptr;
        B2F_CUDA(cudaMallocAsync(&q, sizeof(T) * (size_t)total, st));
        tmp.push_back(q);
        ptr = (T *)q;
        return 0;
    };
    const bool fill_on = style == B2F_FILL;
    // one chain of axis passes per requested reduction; the last pass writes the destination (or a full-size temporary when
    // the output box is a crop: Inner(), out axes smaller than the image)
    for (int mode = 0; mode < 2; ++mode) {                   // 0 min, 1 max
        T *dst_final = mode == 0 ? (T *)d_min : (T *)(interleaved ? d_min : d_max);
        if (mode == 0 && !d_min) continue;
        if (mode == 1 && !interleaved && !d_max) continue;
        const long long es = interleaved ? 2 : 1, off = interleaved ? mode : 0;
        const T *src = (const T *)d_img;
        T *ping = nullptr, *pong = nullptr;
        const int np = (int)axes.size();
        if (np == 0) {                                        // 1-element window: a copy
            int rc = vh_pass<T>(src, full ? dst_final : (temp(ping), ping), full ? es : 1, full ? off : 0, dims, nd, 0, 0, 1, mode == 1,
                                false, 0.0, st);
            if (rc) return rc;
            src = full ? nullptr : ping;
        }
        for (int a = 0; a < np; ++a) {
            const int d = axes[a];
            const bool last = a == np - 1;
            T *dst;
            long long des = 1, doff = 0;
            if (last && full) { dst = dst_final; des = es; doff = off; }
            else {
                T *&buf = (a & 1) ? pong : ping;
                if (!buf) { int rc = temp(buf); if (rc) return rc; }
                dst = buf;
            }
            int rc = vh_pass<T>(src, dst, des, doff, dims, nd, d, (int)wlo[d], (int)(whi[d] - wlo[d] + 1), mode == 1, fill_on, fill, st);
            if (rc) return rc;
            src = dst;
        }
        if (!full) {
            const long long o[4] = {out_ax.len(0), out_ax.len(1), out_ax.len(2), out_ax.len(3)};
            const long long f[4] = {out_ax.lo[0] - ia.lo[0], out_ax.lo[1] - ia.lo[1], out_ax.lo[2] - ia.lo[2], out_ax.lo[3] - ia.lo[3]};
            const long long n = o[0] * o[1] * o[2] * o[3];
            const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)sm_count() * 16);
            vh_crop_kernel<T><<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(src, dst_final, es, off, dims[0], dims[1], dims[2], o[0], o[1], o[2], o[3],
                                                                          f[0], f[1], f[2], f[3]);
            count_launch();
            B2F_CUDA(cudaGetLastError());
        }
    }
    return 0;
}

// windows must contain their centre along every axis (truncation == "the in-image part of the window", which is never
// empty then); other windows stay on the generic kernel
bool extrema_vh_applicable(const b2f_array *img, const int64_t *wlo, const int64_t *whi) {
    for (int d = 0; d < img->ndim; ++d) {
        if (wlo[d] > 0 || whi[d] < 0) return false;
        if (whi[d] - wlo[d] + 1 > 4096) return false;
        if (img->dims[d] >= (1LL << 31)) return false;
    }
    return true;
}

int run_extrema_vh(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved, const Box &out_ax,
                   const int64_t *wlo, const int64_t *whi, int style, double fill, cudaStream_t st) {
    set_path("extrema_vh");
#define B2F_EXT(T) return run_extrema_vh_typed<T>(img, d_img, d_min, d_max, interleaved, out_ax, wlo, whi, style, fill, st)
    switch (img->dtype) {
        case B2F_F32: B2F_EXT(float);
        case B2F_F64: B2F_EXT(double);
        case B2F_U8: case B2F_N0F8: B2F_EXT(uint8_t);
        case B2F_I16: B2F_EXT(int16_t);
        case B2F_U16: B2F_EXT(uint16_t);
        case B2F_I32: B2F_EXT(int32_t);
        case B2F_U32: B2F_EXT(uint32_t);
        case B2F_I64: B2F_EXT(long long);
    }
#undef B2F_EXT
    return fail(B2F_EARG, "unsupported dtype");
}

}  // namespace b2f
