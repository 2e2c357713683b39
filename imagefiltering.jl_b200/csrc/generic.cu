// generic.cu — the always-correct device path: one launch per cascade stage, intermediates in HBM.
//
// It follows the reference's schedule literally (src/imfilter.jl:385-395,438-446): stage a is a
// VALID filter over region[a]; only the first stage reads the (virtually padded) image, through the
// border index remap; intermediates have eltype(out) like the reference's `tempbuffer`
// (src/imfilter.jl:1317-1329).  Handles every ndim <= 4, dtype, stage kind and cascade shape
// (repeated axes, dense + 1-D mixes, Laplacian, copy kernels), so the fused kernels only need to
// cover the hot configurations.  Inner loops: src/imfilter.jl:650-669 (dense), :724-739 (1-D),
// src/specialty.jl:3-16 (Laplacian).

#include "common.cuh"

namespace b2f {

template <typename CT>
struct GStage {
    const void *src;
    int src_dt;          // dtype of src (image dtype for the first stage, else the dtype code of CT)
    int src_is_image;
    int64_t src_lo[B2F_MAXDIM], src_len[B2F_MAXDIM];
    int style;
    CT fill;
    void *dst;
    int dst_dt;
    int64_t dst_lo[B2F_MAXDIM], dst_len[B2F_MAXDIM];
    int64_t r_lo[B2F_MAXDIM], r_len[B2F_MAXDIM];
    int kind;
    int64_t klo[B2F_MAXDIM], klen[B2F_MAXDIM];
    const CT *taps;
    int *inexact;
    int check_range;
    long long rlo, rhi;
};

template <typename CT>
__device__ __forceinline__ CT gfetch(const GStage<CT> &g, const int64_t *idx) {
    int64_t off = 0, stride = 1;
#pragma unroll
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        int64_t p = idx[d] - g.src_lo[d];
        if (g.src_is_image) {
            p = remap_index(g.style, p, g.src_len[d]);
            if (p < 0) return g.fill;
        }
        off += p * stride;
        stride *= g.src_len[d];
    }
    return load_elem<CT>(g.src, g.src_dt, off);
}

template <typename CT>
__global__ void __launch_bounds__(256) generic_stage_kernel(const GStage<CT> g) {
    const int64_t total = g.r_len[0] * g.r_len[1] * g.r_len[2] * g.r_len[3];
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t I[B2F_MAXDIM];
        int64_t r = t;
#pragma unroll
        for (int d = 0; d < B2F_MAXDIM; ++d) { I[d] = g.r_lo[d] + r % g.r_len[d]; r /= g.r_len[d]; }
        CT acc;
        if (g.kind == B2F_STAGE_LAPLACIAN) {
            int nfl = 0;
#pragma unroll
            for (int d = 0; d < B2F_MAXDIM; ++d) nfl += (g.klen[d] == 3);
            acc = mul_rn<CT>((CT)(-2 * nfl), gfetch(g, I));
#pragma unroll
            for (int d = 0; d < B2F_MAXDIM; ++d) {
                if (g.klen[d] != 3) continue;
                int64_t J[B2F_MAXDIM] = {I[0], I[1], I[2], I[3]};
                J[d] = I[d] + 1;
                acc = add_rn<CT>(acc, gfetch(g, J));
                J[d] = I[d] - 1;
                acc = add_rn<CT>(acc, gfetch(g, J));
            }
        } else {
            acc = (CT)0;
            int64_t tap = 0;
            for (int64_t j3 = 0; j3 < g.klen[3]; ++j3)
            for (int64_t j2 = 0; j2 < g.klen[2]; ++j2)
            for (int64_t j1 = 0; j1 < g.klen[1]; ++j1)
            for (int64_t j0 = 0; j0 < g.klen[0]; ++j0, ++tap) {
                int64_t J[B2F_MAXDIM] = {I[0] + g.klo[0] + j0, I[1] + g.klo[1] + j1, I[2] + g.klo[2] + j2,
                                         I[3] + g.klo[3] + j3};
                acc = mac<CT>(acc, gfetch(g, J), g.taps[tap]);
            }
        }
        int64_t off = 0, stride = 1;
#pragma unroll
        for (int d = 0; d < B2F_MAXDIM; ++d) { off += (I[d] - g.dst_lo[d]) * stride; stride *= g.dst_len[d]; }
        bool ok = store_elem<CT>(g.dst, g.dst_dt, off, acc);
        if (g.check_range && ok) ok = ((long long)acc >= g.rlo && (long long)acc <= g.rhi);
        if (!ok) atomicExch(g.inexact, 1);
    }
}

template <typename CT> struct DtCode;
template <> struct DtCode<float> { static const int v = B2F_F32; };
template <> struct DtCode<double> { static const int v = B2F_F64; };
template <> struct DtCode<long long> { static const int v = B2F_I64; };

template <typename CT>
static int run_generic_typed(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    const int na = (int)P.active.size();
    int64_t rlo = 0, rhi = 0;
    const bool int_out = int_range(out_dt, rlo, rhi);
    int *d_flag = nullptr;
    AsyncFrees to_free(st);                  // released on every exit, error returns included
    if (int_out) {
        B2F_CUDA(cudaMallocAsync((void **)&d_flag, sizeof(int), st));
        to_free.push_back(d_flag);
        B2F_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    }
    const void *src = d_img;
    int src_dt = img_dt;
    Box src_ax = P.img_ax;
    bool src_is_image = true;
    const int nlaunch = na == 0 ? 1 : na;
    for (int a = 0; a < nlaunch; ++a) {
        GStage<CT> g;
        memset(&g, 0, sizeof g);
        g.src = src; g.src_dt = src_dt; g.src_is_image = src_is_image ? 1 : 0;
        g.style = P.style; g.fill = (CT)P.fill;
        g.inexact = d_flag; g.check_range = int_out ? 1 : 0; g.rlo = rlo; g.rhi = rhi;
        const bool last = (a == nlaunch - 1);
        Box reg = na == 0 ? P.roi : P.region[a];
        CT *d_taps = nullptr;
        if (na == 0) {  // trivial kernel: copyto!(out, R, A, R)  (src/imfilter.jl:372-375)
            g.kind = B2F_STAGE_DENSE;
            for (int d = 0; d < B2F_MAXDIM; ++d) { g.klo[d] = 0; g.klen[d] = 1; }
            CT one = (CT)1;
            B2F_CUDA(cudaMallocAsync((void **)&d_taps, sizeof(CT), st));
            to_free.push_back(d_taps);
            B2F_CUDA(cudaMemcpyAsync(d_taps, &one, sizeof(CT), cudaMemcpyHostToDevice, st));
        } else {
            const StageInfo &si = P.stages[P.active[a]];
            g.kind = si.s->kind == B2F_STAGE_LAPLACIAN ? B2F_STAGE_LAPLACIAN : B2F_STAGE_DENSE;
            int64_t nt = 1;
            for (int d = 0; d < B2F_MAXDIM; ++d) { g.klo[d] = si.lo[d]; g.klen[d] = si.hi[d] - si.lo[d] + 1; nt *= g.klen[d]; }
            if (g.kind != B2F_STAGE_LAPLACIAN) {
                std::vector<CT> h(nt);
                for (int64_t t = 0; t < nt; ++t) h[t] = (CT)si.s->taps[t];
                B2F_CUDA(cudaMallocAsync((void **)&d_taps, sizeof(CT) * nt, st));
                to_free.push_back(d_taps);
                // pageable source: the runtime stages it before returning, so `h` may die here
                B2F_CUDA(cudaMemcpyAsync(d_taps, h.data(), sizeof(CT) * nt, cudaMemcpyHostToDevice, st));
            }
        }
        g.taps = d_taps;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            g.src_lo[d] = src_ax.lo[d]; g.src_len[d] = src_ax.len(d);
            g.r_lo[d] = reg.lo[d]; g.r_len[d] = reg.len(d);
        }
        void *dst;
        Box dst_ax;
        if (last) {
            dst = d_out; dst_ax = P.out_ax; g.dst_dt = out_dt;
        } else {
            B2F_CUDA(cudaMallocAsync(&dst, sizeof(CT) * (size_t)reg.count(), st));
            to_free.push_back(dst);
            dst_ax = reg; g.dst_dt = DtCode<CT>::v;
        }
        g.dst = dst;
        for (int d = 0; d < B2F_MAXDIM; ++d) { g.dst_lo[d] = dst_ax.lo[d]; g.dst_len[d] = dst_ax.len(d); }
        const int64_t total = reg.count();
        int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
        if (blocks < 1) blocks = 1;
        generic_stage_kernel<CT><<<blocks, 256, 0, st>>>(g);
        count_launch();
        B2F_CUDA(cudaGetLastError());
        src = dst; src_dt = DtCode<CT>::v; src_ax = dst_ax; src_is_image = false;
    }
    int rc = 0;
    if (int_out) {
        int flag = 0;
        cudaError_t e = cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(B2F_ECUDA, "flag readback failed: %s", cudaGetErrorString(e));
        else if (flag) rc = fail(B2F_EINEXACT, "result not representable in eltype(out) (InexactError)");
    }
    return rc;
}

int run_generic(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    set_path("generic");
    if (out_dt == B2F_F64) return run_generic_typed<double>(P, d_img, img_dt, d_out, out_dt, st);
    if (out_dt == B2F_F32) return run_generic_typed<float>(P, d_img, img_dt, d_out, out_dt, st);
    return run_generic_typed<long long>(P, d_img, img_dt, d_out, out_dt, st);
}

// ---- generic running extrema: direct N-d window scan (fallback; the tiled kernel is in extrema.cu) ---
template <typename T>
struct GExt {
    const T *img;
    T *omn, *omx;
    int interleaved;
    int64_t len[B2F_MAXDIM];          // image dims
    int64_t o_off[B2F_MAXDIM], o_len[B2F_MAXDIM];  // output box: offset inside the image, dims
    int64_t wlo[B2F_MAXDIM], whi[B2F_MAXDIM];
    int fill_on;
    T fill;
};

template <typename T>
__global__ void __launch_bounds__(256) generic_extrema_kernel(const GExt<T> g) {
    const int64_t total = g.o_len[0] * g.o_len[1] * g.o_len[2] * g.o_len[3];
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t I[B2F_MAXDIM];
        int64_t r = t;
#pragma unroll
        for (int d = 0; d < B2F_MAXDIM; ++d) { I[d] = g.o_off[d] + r % g.o_len[d]; r /= g.o_len[d]; }
        bool have = false, outside = false;
        T mn = T(), mx = T();
        int64_t a[B2F_MAXDIM], b[B2F_MAXDIM];
#pragma unroll
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            a[d] = I[d] + g.wlo[d]; b[d] = I[d] + g.whi[d];
            if (a[d] < 0) { a[d] = 0; outside = true; }
            if (b[d] > g.len[d] - 1) { b[d] = g.len[d] - 1; outside = true; }
        }
        for (int64_t j3 = a[3]; j3 <= b[3]; ++j3)
        for (int64_t j2 = a[2]; j2 <= b[2]; ++j2)
        for (int64_t j1 = a[1]; j1 <= b[1]; ++j1)
        for (int64_t j0 = a[0]; j0 <= b[0]; ++j0) {
            const T v = g.img[j0 + g.len[0] * (j1 + g.len[1] * (j2 + g.len[2] * j3))];
            if (!have) { mn = mx = v; have = true; }
            else { if (v < mn) mn = v; if (v > mx) mx = v; }
        }
        if (g.fill_on && outside) {
            if (!have) { mn = mx = g.fill; }
            else { if (g.fill < mn) mn = g.fill; if (g.fill > mx) mx = g.fill; }
        }
        if (g.interleaved) { g.omn[2 * t] = mn; g.omn[2 * t + 1] = mx; }
        else { if (g.omn) g.omn[t] = mn; if (g.omx) g.omx[t] = mx; }
    }
}

template <typename T>
static int run_extrema_generic_typed(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                                     const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                                     cudaStream_t st) {
    GExt<T> g;
    memset(&g, 0, sizeof g);
    g.img = (const T *)d_img; g.omn = (T *)d_min; g.omx = (T *)d_max; g.interleaved = interleaved;
    Box ia = axes_of(img);
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        g.len[d] = ia.len(d);
        g.o_off[d] = out_ax.lo[d] - ia.lo[d];
        g.o_len[d] = out_ax.len(d);
        g.wlo[d] = d < img->ndim ? wlo[d] : 0;
        g.whi[d] = d < img->ndim ? whi[d] : 0;
    }
    g.fill_on = style == B2F_FILL;
    g.fill = (T)fill;
    const int64_t total = out_ax.count();
    int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    generic_extrema_kernel<T><<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(g);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

int run_extrema_generic(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                        const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                        cudaStream_t st) {
    set_path("extrema_generic");
#define B2F_EXT(T) return run_extrema_generic_typed<T>(img, d_img, d_min, d_max, interleaved, out_ax, wlo, whi, style, fill, st)
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
