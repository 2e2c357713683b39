// fused2d.cuh — K1/K3: fused 2-D separable FIR cascade (+ multi-plane gradients) in ONE launch.
//
// Replaces, for a cascade of two 1-D factors on axes 0 and 1 (KernelFactors.* tuples, imgradients):
//   padarray gather copy                 reference src/border.jl:324-347
//   cascade scheduler + temporaries      reference src/imfilter.jl:385-395,438-446,1317-1323
//   the 1-D inner loop, twice            reference src/imfilter.jl:724-739
//   N independent calls of imgradients   reference src/specialty.jl:47-51 (input read once here)
//
// Semantics are exactly the reference's "pad the input once, then run valid stages over shrinking
// regions": a CTA loads its input tile WITH the halo of the whole cascade, applying the border
// index remap (src/border.jl:564-590) to the global indices at load time, then runs stage 1 over
// tile+halo rows/cols and stage 2 over the tile, both out of shared memory.  Each pixel is read
// from HBM once (halo re-reads hit L2) and written once.
//
// Arithmetic: CT=double -> separate __dmul_rn/__dadd_rn in tap order (bit-exact vs the oracle);
//             CT=float  -> fmaf.
// Thread mapping: the pass along x gives every lane its own ROW and R consecutive outputs kept in a
// register sliding window (128-bit conflict-free LDS thanks to the row pitch); the pass along y gives
// every lane its own COLUMN.  Tap counts are runtime values inside a compile-time bucket LB.
#pragma once

#include "common.cuh"

namespace b2f {

constexpr int F2_MAXTAPS = 32;
constexpr int F2_THREADS = 256;

template <typename CT, int NPL>
struct F2Params {
    const void *img;
    int img_dt;
    int W, H;                 // image extent along axes 0, 1
    long long img_plane;      // elements per batch slice
    void *out[NPL];
    long long out_pitch, out_plane;
    int out_ox, out_oy;       // image-relative coordinate of out's first element
    int rx0, ry0, rw, rh;     // computed region, image-relative
    int style;
    CT fill;
    int Lx, Ly, klox, kloy;
    int BX, BY;               // output tile
    int in_rows, in_cols;     // input tile incl. halo
    int P1, P2;               // smem pitches (elements) of the input tile / the intermediate
    int mid_rows;             // rows of the intermediate (per plane)
    CT kx[NPL][F2_MAXTAPS];
    CT ky[NPL][F2_MAXTAPS];
};

template <typename CT> struct Vec;       // 128-bit shared-memory vector of CT
template <> struct Vec<float> { typedef float4 T; static constexpr int N = 4; };
template <> struct Vec<double> { typedef double2 T; static constexpr int N = 2; };

// ---- pass along x: lane <-> row, R consecutive outputs per thread --------------------------------------
// dst(p, row, c) = sum_j src(row, c + j) * k[p][j]      for row < nrows, c < ncols (ncols % R == 0)
template <typename CT, int R, int LB, int NPL, typename Store>
__device__ __forceinline__ void pass_x(const CT *__restrict__ src, int spitch, int nrows, int ncols,
                                       const CT (*k)[F2_MAXTAPS], int L, Store store) {
    constexpr int VN = Vec<CT>::N;
    constexpr int WIN = ((R + LB - 1 + VN - 1) / VN) * VN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nseg = ncols / R;
    const int ngroups = (nrows + 31) >> 5;
    for (int item = warp; item < nseg * ngroups; item += nwarps) {
        const int seg = item % nseg, grp = item / nseg;
        const int row = (grp << 5) + lane;
        if (row >= nrows) continue;
        CT v[WIN];
        const CT *s = src + (size_t)row * spitch + seg * R;
#pragma unroll
        for (int i = 0; i < WIN; i += VN) {
            if (i < R + L - 1) {
                typename Vec<CT>::T t = *reinterpret_cast<const typename Vec<CT>::T *>(s + i);
                if (VN == 4) { v[i] = ((CT *)&t)[0]; v[i + 1] = ((CT *)&t)[1]; v[i + 2] = ((CT *)&t)[2]; v[i + 3] = ((CT *)&t)[3]; }
                else { v[i] = ((CT *)&t)[0]; v[i + 1] = ((CT *)&t)[1]; }
            }
        }
        CT acc[NPL][R];
#pragma unroll
        for (int p = 0; p < NPL; ++p)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[p][r] = (CT)0;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            if (j < L) {
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    const CT kj = k[p][j];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[p][r] = mac<CT>(acc[p][r], v[r + j], kj);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NPL; ++p) store(p, row, seg * R, acc[p]);
    }
}

// ---- pass along y: lane <-> column, R consecutive output rows per thread -----------------------------------
// dst(p, r, col) = sum_j src[p](r + j, col) * k[p][j]      for r < nrows, col < ncols
template <typename CT, int R, int LB, int NPL, bool SRC_PER_PLANE, typename Store>
__device__ __forceinline__ void pass_y(const CT *__restrict__ src, int spitch, size_t splane, int nrows, int ncols,
                                       const CT (*k)[F2_MAXTAPS], int L, Store store) {
    constexpr int WIN = R + LB - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int nseg = (nrows + R - 1) / R;
    const int ngroups = (ncols + 31) >> 5;
    for (int item = warp; item < nseg * ngroups; item += nwarps) {
        const int grp = item % ngroups, seg = item / ngroups;
        const int col = (grp << 5) + lane;
        if (col >= ncols) continue;
        const int r0 = seg * R;
        if (SRC_PER_PLANE) {
#pragma unroll
            for (int p = 0; p < NPL; ++p) {
                CT v[WIN];
                const CT *s = src + p * splane + (size_t)r0 * spitch + col;
#pragma unroll
                for (int i = 0; i < WIN; ++i)
                    if (i < R + L - 1) v[i] = s[(size_t)i * spitch];
                CT acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = (CT)0;
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    if (j < L) {
                        const CT kj = k[p][j];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = mac<CT>(acc[r], v[r + j], kj);
                    }
                }
                store(p, r0, col, acc);
            }
        } else {
            CT v[WIN];
            const CT *s = src + (size_t)r0 * spitch + col;
#pragma unroll
            for (int i = 0; i < WIN; ++i)
                if (i < R + L - 1) v[i] = s[(size_t)i * spitch];
#pragma unroll
            for (int p = 0; p < NPL; ++p) {
                CT acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = (CT)0;
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    if (j < L) {
                        const CT kj = k[p][j];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = mac<CT>(acc[r], v[r + j], kj);
                    }
                }
                store(p, r0, col, acc);
            }
        }
    }
}

template <typename CT> struct F2R;   // outputs per thread along the filtered axis
template <> struct F2R<float> { static constexpr int RX = 8, RY = 8; };
template <> struct F2R<double> { static constexpr int RX = 4, RY = 4; };

template <typename CT, int LB, int NPL, bool XFIRST>
__global__ void __launch_bounds__(F2_THREADS) fused2d_kernel(const F2Params<CT, NPL> P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int RX = F2R<CT>::RX, RY = F2R<CT>::RY;
    CT *s_in = reinterpret_cast<CT *>(smem_raw);
    CT *s_mid = s_in + (size_t)P.in_rows * P.P1;
    int *s_ix = reinterpret_cast<int *>(s_mid + (size_t)NPL * P.mid_rows * P.P2);
    int *s_iy = s_ix + P.in_cols;

    const int x0 = P.rx0 + blockIdx.x * P.BX;     // first output column of the tile (image-relative)
    const int y0 = P.ry0 + blockIdx.y * P.BY;
    const long long bz = blockIdx.z;

    // border remap tables for this tile (src/border.jl:564-590): -1 = Fill
    for (int c = threadIdx.x; c < P.in_cols; c += blockDim.x)
        s_ix[c] = (int)remap_index(P.style, (int64_t)x0 + P.klox + c, P.W);
    for (int r = threadIdx.x; r < P.in_rows; r += blockDim.x)
        s_iy[r] = (int)remap_index(P.style, (int64_t)y0 + P.kloy + r, P.H);
    __syncthreads();

    // load tile + halo, converting to CT (this is where the reference's padarray converts: src/border.jl:343)
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        const long long base = bz * P.img_plane;
        for (int r = warp; r < P.in_rows; r += nwarps) {
            const int gy = s_iy[r];
            CT *dst = s_in + (size_t)r * P.P1;
            const long long rowoff = base + (long long)gy * P.W;
#pragma unroll 4
            for (int c = lane; c < P.in_cols; c += 32) {
                const int gx = s_ix[c];
                CT v = P.fill;
                if (gx >= 0 && gy >= 0) v = load_elem<CT>(P.img, P.img_dt, rowoff + gx);
                dst[c] = v;
            }
        }
    }
    __syncthreads();

    const int tw = min(P.BX, P.rx0 + P.rw - x0);   // live output columns / rows of this tile
    const int th = min(P.BY, P.ry0 + P.rh - y0);
    auto store_global_rowseg = [&](int p, int row, int c0, const CT *acc) {   // from pass_x: R consecutive x
        if (row >= th) return;
        CT *o = (CT *)P.out[p] + bz * P.out_plane + (long long)(y0 + row - P.out_oy) * P.out_pitch + (x0 + c0 - P.out_ox);
#pragma unroll
        for (int r = 0; r < RX; ++r)
            if (c0 + r < tw) o[r] = acc[r];
    };
    auto store_global_colseg = [&](int p, int r0, int col, const CT *acc) {   // from pass_y: R consecutive y
        if (col >= tw) return;
        CT *o = (CT *)P.out[p] + bz * P.out_plane + (long long)(y0 + r0 - P.out_oy) * P.out_pitch + (x0 + col - P.out_ox);
#pragma unroll
        for (int r = 0; r < RY; ++r)
            if (r0 + r < th) o[(long long)r * P.out_pitch] = acc[r];
    };
    const size_t mid_plane = (size_t)P.mid_rows * P.P2;

    if (XFIRST) {
        // stage 1 along x over every input row (tile + y halo); stage 2 along y over the tile
        auto store_mid = [&](int p, int row, int c0, const CT *acc) {
            CT *d = s_mid + p * mid_plane + (size_t)row * P.P2 + c0;
#pragma unroll
            for (int r = 0; r < RX; r += Vec<CT>::N) {
                typename Vec<CT>::T t;
#pragma unroll
                for (int q = 0; q < Vec<CT>::N; ++q) ((CT *)&t)[q] = acc[r + q];
                *reinterpret_cast<typename Vec<CT>::T *>(d + r) = t;
            }
        };
        pass_x<CT, RX, LB, NPL>(s_in, P.P1, P.in_rows, P.BX, P.kx, P.Lx, store_mid);
        __syncthreads();
        pass_y<CT, RY, LB, NPL, true>(s_mid, P.P2, mid_plane, th, P.BX, P.ky, P.Ly, store_global_colseg);
    } else {
        // stage 1 along y over every input column (tile + x halo); stage 2 along x over the tile
        auto store_mid = [&](int p, int r0, int col, const CT *acc) {
            CT *d = s_mid + p * mid_plane + (size_t)r0 * P.P2 + col;
#pragma unroll
            for (int r = 0; r < RY; ++r) d[(size_t)r * P.P2] = acc[r];
        };
        pass_y<CT, RY, LB, NPL, false>(s_in, P.P1, 0, P.BY, P.in_cols, P.ky, P.Ly, store_mid);
        __syncthreads();
#pragma unroll
        for (int p = 0; p < NPL; ++p) {
            auto store_p = [&](int, int row, int c0, const CT *acc) { store_global_rowseg(p, row, c0, acc); };
            pass_x<CT, RX, LB, 1>(s_mid + p * mid_plane, P.P2, P.BY, P.BX, &P.kx[p], P.Lx, store_p);
        }
    }
}

// host-side launcher for one (CT, NPL) combination; defined in fused2d_f32.cu / fused2d_f64.cu
template <typename CT, int NPL>
int launch_fused2d(F2Params<CT, NPL> &P, bool xfirst, int nbatch, cudaStream_t st);

template <typename CT, int LB, int NPL, bool XFIRST>
static int launch_one(const F2Params<CT, NPL> &P, int nbatch, size_t smem, cudaStream_t st) {
    auto kern = fused2d_kernel<CT, LB, NPL, XFIRST>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid((P.rw + P.BX - 1) / P.BX, (P.rh + P.BY - 1) / P.BY, nbatch);
    kern<<<grid, F2_THREADS, smem, st>>>(P);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

template <typename CT, int NPL>
int launch_fused2d_impl(F2Params<CT, NPL> &P, bool xfirst, int nbatch, cudaStream_t st) {
    constexpr int RX = F2R<CT>::RX, RY = F2R<CT>::RY;
    constexpr int VN = Vec<CT>::N;
    const int Lmax = P.Lx > P.Ly ? P.Lx : P.Ly;
    const int LB = Lmax <= 4 ? 4 : Lmax <= 8 ? 8 : Lmax <= 16 ? 16 : 32;
    P.BX = sizeof(CT) == 4 ? 128 : 64;
    auto pitch = [&](int cols) {   // 128-bit LDS by 8 lanes on 8 different rows must hit 8 different bank groups
        int p = ((cols + VN - 1) / VN) * VN;
        if (sizeof(CT) == 4) { while (p % 8 != 4) p += 4; } else { while (p % 4 != 2) p += 2; }
        return p;
    };
    const int winx = ((RX + P.Lx - 1 + VN - 1) / VN) * VN;   // columns one x-window touches (128-bit granules)
    if (xfirst) {
        P.in_rows = 64;
        P.BY = P.in_rows - (P.Ly - 1);
        P.in_cols = P.BX + P.Lx - 1;
        P.P1 = pitch(P.BX - RX + winx);      // >= in_cols; window over-read stays inside the row
        P.P2 = pitch(P.BX);
        P.mid_rows = P.in_rows + RY;         // the last y-segment of a tile may read a few rows past the data
    } else {
        P.BY = 32;
        P.in_rows = P.BY + P.Ly - 1;
        P.in_cols = P.BX + P.Lx - 1;
        P.P1 = pitch(P.in_cols);
        P.P2 = pitch(P.BX - RX + winx);
        P.mid_rows = P.BY;
    }
    const size_t smem = sizeof(CT) * ((size_t)P.in_rows * P.P1 + (size_t)NPL * P.mid_rows * P.P2) +
                        sizeof(int) * (size_t)(P.in_cols + P.in_rows);
    if (smem > 227 * 1024) return fail(B2F_ENOTSUP, "fused2d tile does not fit shared memory");
    if ((P.rh + P.BY - 1) / P.BY > 65535 || nbatch > 65535) return fail(B2F_ENOTSUP, "fused2d grid too large");
#define B2F_F2(LBv)                                                                              \
    return xfirst ? launch_one<CT, LBv, NPL, true>(P, nbatch, smem, st)                          \
                  : launch_one<CT, LBv, NPL, false>(P, nbatch, smem, st)
    switch (LB) {
        case 4: B2F_F2(4);
        case 8: B2F_F2(8);
        case 16: B2F_F2(16);
        default: B2F_F2(32);
    }
#undef B2F_F2
}

}  // namespace b2f
