// iir.cu — Triggs-Sdika recursive (IIR) filtering along one axis: b2f_iir (include/b2f.h).
//
// Replaces the reference's _imfilter_dim! / leftborder! / rightborder! (src/imfilter.jl:922-1092).  The recursion is serial
// along the filtered axis and independent across lines, so the parallelism is one THREAD per line:
//   * axis >= 1: the array is (W, H, B) with the lines W*B apart by one element — a warp's 32 lines are contiguous in memory,
//     every load and store of the march is one coalesced 128-byte request;
//   * axis 0: lines are rows; a warp takes 32 rows and walks them in 32-column panels through a padded shared-memory tile
//     (coalesced global traffic on both passes, conflict-free transposed access: lane = row).
// Per element the work is 3 dependent multiply-add pairs (kept as separate multiplies and adds, in the reference's order: the
// Float32 / Float64 results are bit-equal to the CPU restatement), so a line costs ~2 * n * 6 dependent operations: the kernel
// is latency-bound by construction (like the reference's loop) and the roofline that matters is "lines in flight x dependent
// op latency"; the pixel loads of the forward pass run 8 elements ahead of the recurrence to keep DRAM latency out of it.
// HBM traffic: forward pass reads img and writes out, backward pass reads and writes out, the scaling is folded into the
// backward pass: 4 array passes per filtered axis.
#include "common.cuh"

namespace b2f {

template <typename T> struct IirCoef {
    T a1, a2, a3, b1, b2, b3, scale, M[3][3], oma, omb, fill;
    int use_fill, copy;
};

template <typename T> __device__ __forceinline__ T iir_step(T x, T c1, T p1, T c2, T p2, T c3, T p3) {
    T t = x;
    t = add_rn<T>(t, mul_rn<T>(c1, p1));
    t = add_rn<T>(t, mul_rn<T>(c2, p2));
    t = add_rn<T>(t, mul_rn<T>(c3, p3));
    return t;
}

// the right-edge initialisation (Triggs & Sdika Eqs. 14-15; rightborder!, src/imfilter.jl:1057-1084): given the forward values
// u[n-1], u[n-2], u[n-3] and u[n-4] .. it returns the final (unscaled) v[n-1], v[n-2], v[n-3]
template <typename T>
__device__ __forceinline__ void iir_right(const IirCoef<T> &K, T iplus, T u1, T u2, T u3, T &v1, T &v2, T &v3) {
    const T uplus = iplus / K.oma, vplus = uplus / K.omb;
    const T d1 = add_rn<T>(u1, -uplus), d2 = add_rn<T>(u2, -uplus), d3 = add_rn<T>(u3, -uplus);
    T vr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        vr[r] = add_rn<T>(add_rn<T>(add_rn<T>(mul_rn<T>(K.M[r][0], d1), mul_rn<T>(K.M[r][1], d2)), mul_rn<T>(K.M[r][2], d3)), vplus);
    v1 = vr[0];
    v2 = iir_step<T>(u2, K.b1, v1, K.b2, vr[1], K.b3, vr[2]);
    v3 = iir_step<T>(u3, K.b1, v2, K.b2, v1, K.b3, vr[1]);
}

// ---- axis >= 1 (and the generic strided form): one thread per line, element stride `es` -------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) iir_strided_kernel(const void *__restrict__ img, int img_dt, T *out, const IirCoef<T> K,
                                                          long long W, long long H, long long nlines) {
    const long long lid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lid >= nlines) return;
    const long long b = lid / W, w = lid - b * W;
    const long long base = b * H * W + w;
    const long long n = H;
    auto X = [&](long long i) -> T { return load_elem<T>(img, img_dt, base + i * W); };
    T *o = out + base;
    if (K.copy) {
        for (long long i = 0; i < n; ++i) o[i * W] = X(i);
        return;
    }
    // forward pass; the last input is read before anything of this line is written (img may be out)
    const T x0 = X(0), xlast = X(n - 1);
    const T uminus = (K.use_fill ? K.fill : x0) / K.oma;
    T p1, p2, p3;                                  // u[i-1], u[i-2], u[i-3]
    p3 = iir_step<T>(x0, K.a1, uminus, K.a2, uminus, K.a3, uminus);
    p2 = iir_step<T>(X(1), K.a1, p3, K.a2, uminus, K.a3, uminus);
    p1 = iir_step<T>(X(2), K.a1, p2, K.a2, p3, K.a3, uminus);
    o[0] = p3; o[W] = p2; o[2 * W] = p1;
    long long i = 3;
    // the pixel loads run one batch of PF elements ahead of the recurrence (double-buffered in registers): with a few thousand
    // lines there are not enough warps to hide a DRAM round trip per batch behind other warps
    constexpr int PF = 8;
    if (i + PF <= n - 1) {
        T x[PF], xn[PF];
#pragma unroll
        for (int u = 0; u < PF; ++u) x[u] = X(i + u);
        for (; i + PF <= n - 1; i += PF) {
            const bool more = i + 2 * PF <= n - 1;
            if (more) {
#pragma unroll
                for (int u = 0; u < PF; ++u) xn[u] = X(i + PF + u);
            }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const T t = iir_step<T>(x[u], K.a1, p1, K.a2, p2, K.a3, p3);
                p3 = p2; p2 = p1; p1 = t;
                o[(i + u) * W] = t;
            }
            if (more) {
#pragma unroll
                for (int u = 0; u < PF; ++u) x[u] = xn[u];
            }
        }
    }
    for (; i <= n - 2; ++i) {
        const T t = iir_step<T>(X(i), K.a1, p1, K.a2, p2, K.a3, p3);
        p3 = p2; p2 = p1; p1 = t;
        o[i * W] = t;
    }
    const T ulast = iir_step<T>(xlast, K.a1, p1, K.a2, p2, K.a3, p3);        // u[n-1]; p1 = u[n-2], p2 = u[n-3]
    T v1, v2, v3;
    iir_right<T>(K, K.use_fill ? K.fill : xlast, ulast, p1, p2, v1, v2, v3);
    o[(n - 1) * W] = mul_rn<T>(v1, K.scale);
    o[(n - 2) * W] = mul_rn<T>(v2, K.scale);
    o[(n - 3) * W] = mul_rn<T>(v3, K.scale);
    // backward pass (reads this thread's own forward values back), scaling folded in: q1 = v[i+1], q2 = v[i+2], q3 = v[i+3]
    T q1 = v3, q2 = v2, q3 = v1;
    i = n - 4;
    if (i - (PF - 1) >= 0) {
        T u[PF], un[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) u[k] = o[(i - k) * W];
        for (; i - (PF - 1) >= 0; i -= PF) {
            const bool more = i - PF - (PF - 1) >= 0;
            if (more) {
#pragma unroll
                for (int k = 0; k < PF; ++k) un[k] = o[(i - PF - k) * W];
            }
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const T t = iir_step<T>(u[k], K.b1, q1, K.b2, q2, K.b3, q3);
                q3 = q2; q2 = q1; q1 = t;
                o[(i - k) * W] = mul_rn<T>(t, K.scale);
            }
            if (more) {
#pragma unroll
                for (int k = 0; k < PF; ++k) u[k] = un[k];
            }
        }
    }
    for (; i >= 0; --i) {
        const T t = iir_step<T>(o[i * W], K.b1, q1, K.b2, q2, K.b3, q3);
        q3 = q2; q2 = q1; q1 = t;
        o[i * W] = mul_rn<T>(t, K.scale);
    }
}

// ---- axis 0: a warp owns 32 rows and walks them in panels of 32 columns through shared memory ---------------------------------
constexpr int IIR_WPB = 4;                          // warps per block
template <typename T>
__global__ void __launch_bounds__(IIR_WPB * 32) iir_rows_kernel(const void *__restrict__ img, int img_dt, T *out, const IirCoef<T> K,
                                                                long long n, long long nrows) {
    __shared__ T tile_s[IIR_WPB][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T(*tile)[33] = tile_s[warp];
    const long long row0 = ((long long)blockIdx.x * IIR_WPB + warp) * 32;
    if (row0 >= nrows) return;
    const int nr = (int)min(32LL, nrows - row0);
    const bool mine = lane < nr;
    const long long myrow = (row0 + (mine ? lane : 0)) * n;
    // panel transfer: for r in rows: lanes run along the columns (coalesced); tile[c][r]
    auto load_panel = [&](long long c0, bool from_out) {
        const int nc = (int)min(32LL, n - c0);
        for (int r = 0; r < nr; ++r) {
            const long long idx = (row0 + r) * n + c0 + lane;
            if (lane < nc) tile[lane][r] = from_out ? out[idx] : load_elem<T>(img, img_dt, idx);
        }
        __syncwarp();
    };
    // the next panel travels in registers while the current one is filtered (a warp has nothing else to hide the round trip)
    T pre[32];
    auto fetch = [&](long long c0, bool from_out) {
        const int nc = (int)min(32LL, n - c0);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const long long idx = (row0 + r) * n + c0 + lane;
            pre[r] = (T)0;
            if (r < nr && lane < nc) pre[r] = from_out ? out[idx] : load_elem<T>(img, img_dt, idx);
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int r = 0; r < 32; ++r) tile[lane][r] = pre[r];
        __syncwarp();
    };
    auto store_panel = [&](long long c0) {
        __syncwarp();
        const int nc = (int)min(32LL, n - c0);
        for (int r = 0; r < nr; ++r)
            if (lane < nc) out[(row0 + r) * n + c0 + lane] = tile[lane][r];
        __syncwarp();
    };
    if (K.copy) {
        for (long long c0 = 0; c0 < n; c0 += 32) { load_panel(c0, false); store_panel(c0); }
        return;
    }
    // the line's first and last pixel, before anything is written (img may be out)
    const T x0 = load_elem<T>(img, img_dt, myrow), xlast = load_elem<T>(img, img_dt, myrow + n - 1);
    const T uminus = (K.use_fill ? K.fill : x0) / K.oma;
    T p1 = uminus, p2 = uminus, p3 = uminus;        // u[-1] = u[-2] = u[-3] = uminus reproduces leftborder! term by term
    fetch(0, false);
    for (long long c0 = 0; c0 < n; c0 += 32) {
        commit();
        if (c0 + 32 < n) fetch(c0 + 32, false);
        const int nc = (int)min(32LL, n - c0);
        if (mine) {
            for (int c = 0; c < nc; ++c) {
                const T t = iir_step<T>(tile[c][lane], K.a1, p1, K.a2, p2, K.a3, p3);
                if (c0 + c < n - 1) { p3 = p2; p2 = p1; p1 = t; }             // after the loop: p1 = u[n-2], p2 = u[n-3]
                tile[c][lane] = t;
            }
        }
        store_panel(c0);
    }
    T v1 = 0, v2 = 0, v3 = 0;
    if (mine) {
        const T ulast = iir_step<T>(xlast, K.a1, p1, K.a2, p2, K.a3, p3);
        iir_right<T>(K, K.use_fill ? K.fill : xlast, ulast, p1, p2, v1, v2, v3);
    }
    // backward pass over the panels, last first; elements n-1, n-2, n-3 take v1, v2, v3
    T q1 = 0, q2 = 0, q3 = 0;
    fetch(((n - 1) / 32) * 32, true);
    for (long long c0 = ((n - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        commit();
        if (c0 >= 32) fetch(c0 - 32, true);
        const int nc = (int)min(32LL, n - c0);
        if (mine) {
            for (int c = nc - 1; c >= 0; --c) {
                const long long i = c0 + c;
                T t;
                if (i == n - 1) t = v1;
                else if (i == n - 2) t = v2;
                else if (i == n - 3) t = v3;
                else t = iir_step<T>(tile[c][lane], K.b1, q1, K.b2, q2, K.b3, q3);
                q3 = q2; q2 = q1; q1 = t;
                tile[c][lane] = mul_rn<T>(t, K.scale);
            }
        }
        store_panel(c0);
    }
}

template <typename T>
static int run_iir_typed(const b2f_array *img, const void *d_img, void *d_out, int axis, const double *coef, int style, double fill,
                         cudaStream_t st) {
    IirCoef<T> K;
    K.a1 = (T)coef[0]; K.a2 = (T)coef[1]; K.a3 = (T)coef[2]; K.b1 = (T)coef[3]; K.b2 = (T)coef[4]; K.b3 = (T)coef[5];
    K.scale = (T)coef[6];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) K.M[r][c] = (T)coef[7 + 3 * r + c];
    K.oma = (T)coef[16]; K.omb = (T)coef[17];
    K.fill = (T)fill; K.use_fill = style == B2F_FILL;
    K.copy = coef[6] == 1.0;
    for (int i = 0; i < 6; ++i) K.copy = K.copy && coef[i] == 0.0;
    long long W = 1, H = img->dims[axis], B = 1;
    for (int d = 0; d < axis; ++d) W *= img->dims[d];
    for (int d = axis + 1; d < img->ndim; ++d) B *= img->dims[d];
    if (W == 1 && H >= 64) {                        // contiguous lines: the panel kernel
        const long long nrows = B, blocks = (nrows + 32 * IIR_WPB - 1) / (32 * IIR_WPB);
        if (blocks >= (1LL << 31)) return fail(B2F_ENOTSUP, "array too large for one IIR launch");
        set_path("iir_rows");
        iir_rows_kernel<T><<<(unsigned)blocks, IIR_WPB * 32, 0, st>>>(d_img, img->dtype, (T *)d_out, K, H, nrows);
    } else {
        const long long nlines = W * B, blocks = (nlines + 127) / 128;
        if (blocks >= (1LL << 31)) return fail(B2F_ENOTSUP, "array too large for one IIR launch");
        set_path("iir_strided");
        iir_strided_kernel<T><<<(unsigned)blocks, 128, 0, st>>>(d_img, img->dtype, (T *)d_out, K, W, H, nlines);
    }
    count_launch(1);
    B2F_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_iir(const b2f_array *img, const b2f_array *out, int32_t axis, const double *coef, const b2f_border *border,
                       void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !out || !coef || !border) return fail(B2F_EARG, "NULL argument");
    const int N = img->ndim;
    if (N < 1 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "IIR filtering needs 1..4 dims and equal rank");
    if (axis < 0 || axis >= N) return fail(B2F_EARG, "axis %d outside the array", (int)axis);
    int64_t total = 1;
    for (int d = 0; d < N; ++d) {
        if (img->dims[d] != out->dims[d]) return fail(B2F_EDIM, "out must have the axes of img");
        total *= img->dims[d] < 0 ? 0 : img->dims[d];
    }
    if (border->style != B2F_REPLICATE && border->style != B2F_FILL) return fail(B2F_EARG, "only \"replicate\" is supported");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "IIR filtering produces Float32 / Float64 arrays");
    if (total == 0) { set_path("empty"); return 0; }
    bool copy = coef[6] == 1.0;
    for (int i = 0; i < 6; ++i) copy = copy && coef[i] == 0.0;
    if (!copy && img->dims[axis] <= 3)
        return fail(B2F_EDIM, "size %lld of img along dimension %d is too small for filtering with IIR kernel of length 3",
                    (long long)img->dims[axis], (int)axis + 1);
    int rc = ensure_ctx();
    if (rc) return rc;
    // host arrays are staged like b2f_imfilter's; an in-place call (out is img) stages once
    const bool inplace = img->ptr == out->ptr && img->dtype == out->dtype;
    Staged sin, sout;
    rc = stage_in(img, sin, st, true);
    if (!rc && !inplace) rc = stage_in(out, sout, st, false);
    if (!rc) {
        void *d_out = inplace ? sin.dptr : sout.dptr;
        rc = out->dtype == B2F_F32 ? run_iir_typed<float>(img, sin.dptr, d_out, axis, coef, border->style, border->fill, st)
                                   : run_iir_typed<double>(img, sin.dptr, d_out, axis, coef, border->style, border->fill, st);
        if (!rc && out->mem == B2F_HOST) {
            const Staged &so = inplace ? sin : sout;
            cudaError_t e = cudaMemcpyAsync(out->ptr, d_out, so.bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = fail(B2F_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
        }
    }
    release(sin, st);
    if (!inplace) release(sout, st);
    if (img->mem == B2F_HOST || out->mem == B2F_HOST) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && !rc) rc = fail(B2F_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    return rc;
}
