// extrema2d.cuh — K4: running min / max / extrema over a (wx, wy) window of a batch of 2-D images.
//
// Replaces the reference's Lemire streaming max-min + permutedims per axis (src/mapwindow.jl:388-481,
// `mapwindow(extrema, A, window)`) and the O(prod(w)) generic window copy that `minimum`/`maximum` take
// (src/mapwindow.jl:270-333).  min/max are exact in any evaluation order, so the result is bit-identical
// for NaN-free data.  Window placement [i+lo, i+hi] and truncation at the array ends follow
// src/mapwindow.jl:426-473 (even widths) / :136-150 (odd Dims, ranges); Fill adds the fill value where the
// window leaves the array (:326-333).  For windows that contain their centre, truncation == replicate
// index remapping, which is how the loader implements it.
//
// Same warp-streamed organisation as stream2d.cuh: a warp owns 128 columns x SH rows, prefetches RB
// rows ahead, does the x-window from a register sliding window and keeps the y-window as
// output-stationary running extrema in a register ring that rotates at compile time.
#pragma once

#include <cfloat>

#include "common.cuh"
#include "stream2d.cuh"

namespace b2f {

enum { EX_MIN = 0, EX_MAX = 1, EX_BOTH = 2, EX_PAIR = 3 };   // PAIR: interleaved (min,max) tuples

struct E2Params {
    const void *img;
    int W, H;
    long long img_plane;
    void *omin, *omax;        // PAIR: omin holds the tuples
    long long out_pitch, out_plane;   // in elements (tuples for PAIR)
    int out_ox, out_oy;
    int rx0, ry0, rw, rh;
    int style;                // B2F_REPLICATE (truncate) or B2F_FILL
    float fill;
    int Wx, Wy, lox, loy;
    int SH, nsx, nsy;
    long long nstrips;
    int vec_ok;
};

template <int WXT, int WYT, int LB, int MODE, int RB, int ROT>
__device__ __forceinline__ void e2_row(const int u, const int rv, const E2Params &P, const int Wx, const int Wy,
                                       const int th, const float *__restrict__ sblk, const int lane,
                                       const int tw, const bool lane_full, const bool lane_live,
                                       float (&amn)[ROT][4],
                                       float (&amx)[ROT][4],
                                       float *&pmn, float *&pmx) {
    constexpr int PX = 4;
    constexpr int LBX = WXT ? WXT : LB;
    constexpr int LBY = WYT ? WYT : LB;
    static_assert(ROT >= LBY, "ring shorter than the y window");
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX;
    constexpr int PW = 32 * PX + WIN;
    constexpr bool DO_MIN = MODE != EX_MAX, DO_MAX = MODE != EX_MIN;
    const float *srow = sblk + (u % RB) * PW + lane * PX;
    float v[WIN];
#pragma unroll
    for (int i = 0; i < WIN; i += PX) {
        if (WXT || i < PX + Wx - 1) {
            float4 t = *reinterpret_cast<const float4 *>(srow + i);
            v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
        }
    }
    float xmn[PX], xmx[PX];
#pragma unroll
    for (int q = 0; q < PX; ++q) { xmn[q] = v[q]; xmx[q] = v[q]; }
#pragma unroll
    for (int j = 1; j < LBX; ++j) {
        if (WXT || j < Wx) {
#pragma unroll
            for (int q = 0; q < PX; ++q) {
                if (DO_MIN) xmn[q] = fminf(xmn[q], v[q + j]);
                if (DO_MAX) xmx[q] = fmaxf(xmx[q], v[q + j]);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < LBY; ++d) {
        if (WYT || d < Wy) {
            const int slot = (u + 1 + d) % ROT;
#pragma unroll
            for (int q = 0; q < PX; ++q) {
                if (DO_MIN) amn[slot][q] = fminf(amn[slot][q], xmn[q]);
                if (DO_MAX) amx[slot][q] = fmaxf(amx[slot][q], xmx[q]);
            }
        }
    }
    const int eslot = (u + 1) % ROT;
    const int o = rv - (ROT - 1);
    if (o >= 0 && o < th) {
        if (MODE == EX_PAIR) {
            if (lane_full) {
                float4 a = make_float4(amn[eslot][0], amx[eslot][0], amn[eslot][1], amx[eslot][1]);
                float4 b = make_float4(amn[eslot][2], amx[eslot][2], amn[eslot][3], amx[eslot][3]);
                reinterpret_cast<float4 *>(pmn)[0] = a;
                reinterpret_cast<float4 *>(pmn)[1] = b;
            } else if (lane_live) {
#pragma unroll
                for (int q = 0; q < PX; ++q)
                    if (lane * PX + q < tw) { pmn[2 * q] = amn[eslot][q]; pmn[2 * q + 1] = amx[eslot][q]; }
            }
            pmn += 2 * P.out_pitch;
        } else {
            if (lane_full) {
                if (DO_MIN) *reinterpret_cast<float4 *>(pmn) = make_float4(amn[eslot][0], amn[eslot][1], amn[eslot][2], amn[eslot][3]);
                if (DO_MAX) *reinterpret_cast<float4 *>(pmx) = make_float4(amx[eslot][0], amx[eslot][1], amx[eslot][2], amx[eslot][3]);
            } else if (lane_live) {
#pragma unroll
                for (int q = 0; q < PX; ++q)
                    if (lane * PX + q < tw) {
                        if (DO_MIN) pmn[q] = amn[eslot][q];
                        if (DO_MAX) pmx[q] = amx[eslot][q];
                    }
            }
            if (DO_MIN) pmn += P.out_pitch;
            if (DO_MAX) pmx += P.out_pitch;
        }
    }
#pragma unroll
    for (int q = 0; q < PX; ++q) { amn[eslot][q] = FLT_MAX * 2.0f; amx[eslot][q] = -FLT_MAX * 2.0f; }   // +inf / -inf
}

template <int WXT, int WYT, int LB, int MODE, int RB, int ROT>
__global__ void __launch_bounds__(S2_WARPS * 32) extrema2d_kernel(const E2Params P) {
    constexpr int PX = 4;
    constexpr int CW = 32 * PX;
    constexpr int LBX = WXT ? WXT : LB;
    constexpr int LBY = WYT ? WYT : LB;
    constexpr int G = (ROT / s2_gcd(ROT, RB)) * RB;
    constexpr int NCL = (CW + LBX - 1 + 31) / 32;
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX;
    constexpr int PW = CW + WIN;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // shuffle: provably warp-uniform
    float *sbuf = reinterpret_cast<float *>(smem_raw) + (size_t)warp * (2 * RB * PW);

    const long long sid = (long long)blockIdx.x * S2_WARPS + warp;
    if (sid >= P.nstrips) return;
    const int sx = (int)(sid % P.nsx);
    const int sy = (int)((sid / P.nsx) % P.nsy);
    const long long bz = sid / ((long long)P.nsx * P.nsy);

    const int Wx = WXT ? WXT : P.Wx;
    const int Wy = WYT ? WYT : P.Wy;
    const int x0 = P.rx0 + sx * CW;
    const int y0 = P.ry0 + sy * P.SH;
    const int tw = min(CW, P.rx0 + P.rw - x0);
    const int th = min(P.SH, P.ry0 + P.rh - y0);
    const int in_rows = th + Wy - 1;
    const int in_cols = CW + Wx - 1;
    const float *__restrict__ img = reinterpret_cast<const float *>(P.img) + bz * P.img_plane;
    const bool is_fill = P.style == B2F_FILL;

    int gx[NCL];
    unsigned colfill = 0, coldead = 0;
#pragma unroll
    for (int c = 0; c < NCL; ++c) {
        const int col = lane + 32 * c;
        int g = 0;
        if (col < in_cols) {
            g = s2_remap(P.style, x0 + P.lox + col, P.W);
            if (g < 0) { colfill |= 1u << c; g = 0; }
        } else {
            coldead |= 1u << c;
        }
        gx[c] = g;
    }
    const int ytop = y0 + P.loy;
    const bool y_interior = ytop >= 0 && ytop + in_rows <= P.H;
    const int s0 = ROT - Wy;
    const int vrows = in_rows + s0;

    float stage[RB][NCL];
    unsigned rowfill = 0;
    auto fetch_block = [&](int blk) {
        rowfill = 0;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
            int r = blk * RB + rr - s0;
            r = min(max(r, 0), in_rows - 1);
            int gy = ytop + r;
            if (!y_interior) {
                gy = s2_remap(P.style, gy, P.H);
                if (gy < 0) { rowfill |= 1u << rr; gy = 0; }
            }
            const float *row = img + (long long)gy * P.W;
#pragma unroll
            for (int c = 0; c < NCL; ++c) stage[rr][c] = __ldg(row + gx[c]);
        }
    };
    auto park_block = [&](int blk) {
        float *dst = sbuf + (blk & 1) * (RB * PW) + lane;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
            for (int c = 0; c < NCL; ++c) {
                float v = stage[rr][c];
                if (is_fill && ((colfill >> c | rowfill >> rr) & 1u)) v = P.fill;
                if (c < NCL - 1 || !((coldead >> c) & 1u)) dst[rr * PW + 32 * c] = v;
            }
        }
    };

    float amn[ROT][PX], amx[ROT][PX];
#pragma unroll
    for (int s = 0; s < ROT; ++s)
#pragma unroll
        for (int q = 0; q < PX; ++q) { amn[s][q] = FLT_MAX * 2.0f; amx[s][q] = -FLT_MAX * 2.0f; }

    const long long off = bz * P.out_plane + (long long)(y0 - P.out_oy) * P.out_pitch + (x0 + lane * PX - P.out_ox);
    float *pmn = reinterpret_cast<float *>(P.omin) + (MODE == EX_PAIR ? 2 * off : off);
    float *pmx = reinterpret_cast<float *>(P.omax) + off;
    const bool lane_full = P.vec_ok && (lane * PX + PX <= tw);
    const bool lane_live = lane * PX < tw;

    fetch_block(0);
    park_block(0);
    __syncwarp();

    // Every group of ROT rows runs the same unchecked code: rows before the strip / past its end are clamped
    // duplicates whose contributions only reach output rows that are never stored (o < 0 or o >= th).
    for (int rbase = 0; rbase < vrows; rbase += G) {
        const int blk0 = rbase / RB;
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int blk = blk0 + u / RB;
            if (u % RB == 0) fetch_block(blk + 1);                // loads fly while this block is computed
            e2_row<WXT, WYT, LB, MODE, RB, ROT>(u, rbase + u, P, Wx, Wy, th, sbuf + (blk & 1) * (RB * PW), lane, tw, lane_full,
                                          lane_live, amn, amx, pmn, pmx);
            if (u % RB == RB - 1) {
                park_block(blk + 1);
                __syncwarp();
            }
        }
    }
}

template <int WXT, int WYT, int LB, int MODE, int RB, int ROT>
static int e2_launch_one(const E2Params &P, cudaStream_t st) {
    constexpr int PX = 4;
    constexpr int LBX = WXT ? WXT : LB;
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX;
    constexpr int PW = 32 * PX + WIN;
    const size_t smem = (size_t)S2_WARPS * 2 * RB * PW * sizeof(float);
    const long long blocks = (P.nstrips + S2_WARPS - 1) / S2_WARPS;
    if (blocks > 0x7fffffffLL) return fail(B2F_ENOTSUP, "extrema2d grid too large");
    extrema2d_kernel<WXT, WYT, LB, MODE, RB, ROT><<<(unsigned)blocks, S2_WARPS * 32, smem, st>>>(P);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

template <int MODE>
static int e2_launch(const E2Params &P, cudaStream_t st) {
    const int Wx = P.Wx, Wy = P.Wy, L = Wx > Wy ? Wx : Wy;
    if (Wx == 3 && Wy == 3) return e2_launch_one<3, 3, 4, MODE, 6, 3>(P, st);
    if (Wx == 5 && Wy == 5) return e2_launch_one<5, 5, 8, MODE, 3, 6>(P, st);
    if (Wx == 7 && Wy == 7) return e2_launch_one<7, 7, 8, MODE, 4, 8>(P, st);
    if (Wx == 9 && Wy == 9) return e2_launch_one<9, 9, 16, MODE, 3, 9>(P, st);
    if (L <= 4) return e2_launch_one<0, 0, 4, MODE, 4, 4>(P, st);
    if (L <= 8) return e2_launch_one<0, 0, 8, MODE, 4, 8>(P, st);
    if (L <= 16) return e2_launch_one<0, 0, 16, MODE, 4, 16>(P, st);
    return fail(B2F_ENOTSUP, "extrema2d: window outside the instantiated range");
}

int launch_extrema2d_min(const E2Params &P, cudaStream_t st);
int launch_extrema2d_max(const E2Params &P, cudaStream_t st);
int launch_extrema2d_both(const E2Params &P, cudaStream_t st);
int launch_extrema2d_pair(const E2Params &P, cudaStream_t st);

}  // namespace b2f
