// stream2d.cuh — K1/K3 fast path: warp-streamed fused 2-D separable FIR (+ 2-plane gradients).
//
// Same semantics and arithmetic as fused2d.cuh (reference src/imfilter.jl:385-395,438-446,724-739,
// src/border.jl:564-590, src/specialty.jl:47-51) but organised for HBM throughput:
//
//   * every WARP owns a strip of CW = 32*PX columns x SH rows and works alone (no __syncthreads):
//     latency is hidden by many independent warps, not by CTA-wide phases;
//   * the strip is marched along y.  Input rows are fetched RB rows ahead into registers (all
//     loads of a block in flight at once), converted to the compute type and parked in a small
//     per-warp double-buffered shared-memory ring (this is where the border remap is applied);
//   * stage 1 (along x): each lane produces PX adjacent outputs of a row from a register sliding
//     window read with 128-bit conflict-free LDS;
//   * stage 2 (along y): output-stationary accumulators in registers.  Row r adds mid[r]*ky[j] to the
//     LY outputs whose window contains it, in ascending tap order, so the bits match the
//     reference loop; the accumulator ring rotates at compile time (row loop unrolled by LB), so
//     there are no register moves and no shared-memory intermediate;
//   * a finished output row leaves as one 128-bit store per lane and plane (512 B per warp).
//
// HBM traffic: each input pixel is read once (+ halo re-reads that hit L2), each output written once.
#pragma once

#include "common.cuh"

namespace b2f {

constexpr int S2_MAXTAPS = 16;
constexpr int S2_WARPS = 4;   // warps per CTA (they do not cooperate)

template <typename CT, int NPL>
struct S2Params {
    const void *img;
    int n0f8;                 // u8 input means N0f8 (i/255) rather than the integer i
    int W, H;
    long long img_plane;
    void *out[NPL];
    long long out_pitch, out_plane;
    int out_ox, out_oy;
    int rx0, ry0, rw, rh;
    int style;
    CT fill;
    int Lx, Ly, klox, kloy;
    int SH;                   // output rows per strip
    int nsx, nsy;             // strips along x / y
    long long nstrips;        // nsx * nsy * batch
    int vec_ok;               // output rows are 16-byte aligned for every strip
    CT kx[NPL][S2_MAXTAPS];
    CT kyr[NPL][S2_MAXTAPS];  // y taps reversed: kyr[d] = ky[Ly-1-d]
};

template <typename CT> struct S2Vec;
template <> struct S2Vec<float> { typedef float4 T; static constexpr int PX = 4; };
template <> struct S2Vec<double> { typedef double2 T; static constexpr int PX = 2; };

template <typename IT, typename CT> struct S2Conv {
    __device__ static __forceinline__ CT f(IT v, int) { return (CT)v; }
};
template <> struct S2Conv<uint8_t, double> {
    __device__ static __forceinline__ double f(uint8_t v, int n0f8) { return n0f8 ? n0f8_to_f64(v) : (double)v; }
};
template <> struct S2Conv<uint8_t, float> {
    __device__ static __forceinline__ float f(uint8_t v, int n0f8) { return n0f8 ? n0f8_to_f32(v) : (float)v; }
};

template <typename IT, typename CT, int LB, int NPL>
__global__ void __launch_bounds__(S2_WARPS * 32) stream2d_kernel(const S2Params<CT, NPL> P) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int CW = 32 * PX;                         // strip width
    constexpr int RB = 4;                               // rows per prefetch block
    constexpr int G = LB > RB ? LB : RB;                // rows per unrolled outer iteration
    constexpr int NCL = (CW + LB - 1 + 31) / 32;        // loads per lane per input row
    constexpr int WIN = ((PX + LB - 1 + PX - 1) / PX) * PX;   // window registers (whole 128-bit granules)
    constexpr int PW = CW + WIN;                        // smem row pitch (elements), multiple of PX
    typedef typename S2Vec<CT>::T V;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    CT *sbuf = reinterpret_cast<CT *>(smem_raw) + (size_t)warp * (2 * RB * PW);

    const long long sid = (long long)blockIdx.x * S2_WARPS + warp;
    if (sid >= P.nstrips) return;
    const int sx = (int)(sid % P.nsx);
    const int sy = (int)((sid / P.nsx) % P.nsy);
    const long long bz = sid / ((long long)P.nsx * P.nsy);

    const int x0 = P.rx0 + sx * CW;
    const int y0 = P.ry0 + sy * P.SH;
    const int tw = min(CW, P.rx0 + P.rw - x0);          // live output columns
    const int th = min(P.SH, P.ry0 + P.rh - y0);        // live output rows
    const int in_rows = th + P.Ly - 1;
    const int in_cols = CW + P.Lx - 1;
    const IT *__restrict__ img = reinterpret_cast<const IT *>(P.img) + bz * P.img_plane;

    // remapped source column of each of this lane's load slots (-1: Fill, -2: beyond the tile)
    int gx[NCL];
#pragma unroll
    for (int c = 0; c < NCL; ++c) {
        const int col = lane + 32 * c;
        gx[c] = col < in_cols ? (int)remap_index(P.style, (int64_t)x0 + P.klox + col, P.W) : -2;
    }

    // Virtual row index rv = r + s0 with s0 = LB - Ly: then output row o = r - (Ly-1) = rv - (LB-1) always sits in
    // accumulator slot (rv+1) % LB and row r feeds slot (rv+1+d) % LB with tap ky[Ly-1-d] — every register index
    // is a compile-time constant once the row loop is unrolled by LB (P.kyr holds the taps reversed).
    const int s0 = LB - P.Ly;
    const int vrows = in_rows + s0;

    IT stage[RB][NCL];
    int stage_ok[RB];   // bit c set: slot c of that row is a real pixel (else Fill)
    auto fetch_block = [&](int blk) {
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
            const int r = blk * RB + rr - s0;
            int ok = 0;
            if (r >= 0 && r < in_rows) {
                const int gy = (int)remap_index(P.style, (int64_t)y0 + P.kloy + r, P.H);
                const IT *row = img + (long long)gy * P.W;
#pragma unroll
                for (int c = 0; c < NCL; ++c) {
                    if (gy >= 0 && gx[c] >= 0) { stage[rr][c] = __ldg(row + gx[c]); ok |= 1 << c; }
                }
            }
            stage_ok[rr] = ok;
        }
    };
    auto park_block = [&](int blk) {
        CT *dst = sbuf + (size_t)(blk & 1) * (RB * PW);
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
            for (int c = 0; c < NCL; ++c) {
                if (gx[c] != -2) {
                    CT v = P.fill;
                    if (stage_ok[rr] >> c & 1) v = S2Conv<IT, CT>::f(stage[rr][c], P.n0f8);
                    dst[rr * PW + lane + 32 * c] = v;
                }
            }
        }
    };

    CT acc[NPL][LB][PX];
#pragma unroll
    for (int p = 0; p < NPL; ++p)
#pragma unroll
        for (int s = 0; s < LB; ++s)
#pragma unroll
            for (int q = 0; q < PX; ++q) acc[p][s][q] = (CT)0;

    const int nblk = (vrows + RB - 1) / RB;
    fetch_block(0);
    park_block(0);
    __syncwarp();

    const bool lane_live = lane * PX < tw;
    const bool lane_full = P.vec_ok && (lane * PX + PX <= tw);
    for (int rbase = 0; rbase < vrows; rbase += G) {
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int rv = rbase + u;
            const int blk = rv / RB;
            if (u % RB == 0 && blk + 1 < nblk) fetch_block(blk + 1);   // loads fly while this block is computed
            if (rv >= s0 && rv < vrows) {
                // ---- stage 1 along x: PX outputs from a register window -------------------------------
                const CT *srow = sbuf + (size_t)(blk & 1) * (RB * PW) + (u % RB) * PW + lane * PX;
                CT v[WIN];
#pragma unroll
                for (int i = 0; i < WIN; i += PX) {
                    if (i < PX + P.Lx - 1) {
                        V t = *reinterpret_cast<const V *>(srow + i);
#pragma unroll
                        for (int q = 0; q < PX; ++q) v[i + q] = ((CT *)&t)[q];
                    }
                }
                CT mid[NPL][PX];
#pragma unroll
                for (int p = 0; p < NPL; ++p)
#pragma unroll
                    for (int q = 0; q < PX; ++q) mid[p][q] = (CT)0;
#pragma unroll
                for (int j = 0; j < LB; ++j) {
                    if (j < P.Lx) {
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            const CT kj = P.kx[p][j];
#pragma unroll
                            for (int q = 0; q < PX; ++q) mid[p][q] = mac<CT>(mid[p][q], v[q + j], kj);
                        }
                    }
                }
                // ---- stage 2 along y: this row is tap Ly-1-d of the output held in slot (u+1+d) % LB -----------
                // (d descending = the order in which the reference adds taps to each output: ascending j)
#pragma unroll
                for (int d = 0; d < LB; ++d) {
                    if (d < P.Ly) {
                        const int slot = (u + 1 + d) % LB;
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            const CT kj = P.kyr[p][d];
#pragma unroll
                            for (int q = 0; q < PX; ++q) acc[p][slot][q] = mac<CT>(acc[p][slot][q], mid[p][q], kj);
                        }
                    }
                }
                // ---- output row o = rv-(LB-1) is complete: emit it and recycle its slot -----------------------
                {
                    const int slot = (u + 1) % LB;
                    const int o = rv - (LB - 1);
                    if (o >= 0 && lane_live) {
                        const long long off = bz * P.out_plane + (long long)(y0 + o - P.out_oy) * P.out_pitch +
                                              (x0 + lane * PX - P.out_ox);
#pragma unroll
                        for (int p = 0; p < NPL; ++p) {
                            CT *dst = reinterpret_cast<CT *>(P.out[p]) + off;
                            if (lane_full) {
                                V t;
#pragma unroll
                                for (int q = 0; q < PX; ++q) ((CT *)&t)[q] = acc[p][slot][q];
                                *reinterpret_cast<V *>(dst) = t;
                            } else {
#pragma unroll
                                for (int q = 0; q < PX; ++q)
                                    if (lane * PX + q < tw) dst[q] = acc[p][slot][q];
                            }
                        }
                    }
#pragma unroll
                    for (int p = 0; p < NPL; ++p)
#pragma unroll
                        for (int q = 0; q < PX; ++q) acc[p][slot][q] = (CT)0;
                }
            }
            if (u % RB == RB - 1 && blk + 1 < nblk) {
                park_block(blk + 1);
                __syncwarp();
            }
        }
    }
}

}  // namespace b2f
