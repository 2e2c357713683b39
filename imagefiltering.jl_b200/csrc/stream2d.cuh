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
//   * stage 2 (along y) runs in TRANSPOSED (systolic) form, in registers: acc[j] is the partial sum of the output
//     that takes tap j next; a new x-filtered row `mid` completes the oldest output (tap Ly-1: emitted) and moves
//     every other partial sum one slot up while it adds its tap, acc[j+1] = mid*ky[j] + acc[j].  Each output still
//     receives its taps in ascending order, so the bits match the reference loop; the register indices are static
//     without unrolling the row loop by the tap count (that version spread the hot loop over 59 KB of code and
//     stalled on instruction fetch), no register moves, no shared-memory intermediate;
//   * a finished output row leaves as one 128-bit store per lane and plane (512 B per warp).
//
// Tap counts: LXT/LYT > 0 are compile-time exact (hot sizes: no predicates at all); LXT = LYT = 0 means
// run-time counts bounded by the bucket LB (uniform predicates around each tap).
//
// HBM traffic: each input pixel is read once (+ halo re-reads that hit L2), each output written once.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace b2f {

constexpr int S2_MAXTAPS = 20;
constexpr int S2_WARPS = 4;   // warps per CTA (they do not cooperate)

template <typename CT, int NPL>
struct S2Params {
    const void *img;
    CT n0_r, n0_c;            // u8 -> CT: q=x*r; q += fma(-q,c,x)*r.  (1/255,255) for N0f8, (1,1) for raw bytes
    int W, H;                 // extent of the array along x / y (y: rows present in this buffer)
    int Hg, y_first;          // slab form: the full axis has Hg rows and buffer row 0 is global row y_first
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
    int fma;                  // host side only: accum mode B2F_ACCUM_FMA (Float64 compute: fused multiply-add where instantiated)
    CT kx[NPL][S2_MAXTAPS];
    CT ky[NPL][S2_MAXTAPS];   // y taps, ascending (filled by the host set-up)
    CT kyt[NPL][S2_MAXTAPS];  // the same taps RIGHT-aligned in the instantiation's LBY slots (filled by the launcher)
    float2 kxp[NPL][S2_MAXTAPS];  // Float32 compute only: kxp[j] = (kx[j], kx[j-1]), the taps one input carries to two adjacent outputs
};

// Packed FP32 multiply-add (sm_100a FFMA2: two FMAs per issue slot; every result is the same single-rounded fma as fmaf).
// s2_fma2: (a.x*k + c.x, a.y*k + c.y), tap broadcast.  s2_fma2b: (v*k.x + c.x, v*k.y + c.y), value broadcast.
__device__ __forceinline__ float2 s2_fma2(float2 a, float k, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rc = *reinterpret_cast<unsigned long long *>(&c), rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(k));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 s2_fma2b(float v, float2 k, float2 c) {
    unsigned long long rk = *reinterpret_cast<unsigned long long *>(&k), rc = *reinterpret_cast<unsigned long long *>(&c), rv, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rv), "l"(rk), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

template <typename CT> struct S2Vec;
template <> struct S2Vec<float> { typedef float4 T; static constexpr int PX = 4; };
template <> struct S2Vec<double> { typedef double2 T; static constexpr int PX = 2; };

template <typename IT, typename CT> struct S2Conv {
    __device__ static __forceinline__ CT f(IT v, CT, CT) { return (CT)v; }
};
template <> struct S2Conv<uint8_t, double> {   // branch-free, correctly rounded i/255 (or i itself when r=c=1)
    __device__ static __forceinline__ double f(uint8_t v, double r, double c) {
        const double x = (double)v;
        const double q = __dmul_rn(x, r);
        return fma(fma(-q, c, x), r, q);
    }
};
template <> struct S2Conv<uint8_t, float> {
    __device__ static __forceinline__ float f(uint8_t v, float r, float c) {
        const float x = (float)v;
        const float q = __fmul_rn(x, r);
        return fmaf(fmaf(-q, c, x), r, q);
    }
};

constexpr int s2_gcd(int a, int b) { return b == 0 ? a : s2_gcd(b, a % b); }

// 32-bit twin of remap_index.  The out-of-range branch is kept OUT OF LINE: it runs for border strips only, and inlined
// into every unrolled row it spread the hot loop over 79 KB of code (instruction-cache misses were the top stall).
static __device__ __noinline__ int s2_remap_slow(int style, int i, int n) {
    if (style == B2F_REPLICATE) return i < 0 ? 0 : n - 1;
    if (style == B2F_FILL) return -1;
    int p = style == B2F_CIRCULAR ? n : (style == B2F_SYMMETRIC ? 2 * n : 2 * n - 2);
    int m = i % p;
    if (m < 0) m += p;
    if (style == B2F_CIRCULAR || m < n) return m;
    return style == B2F_SYMMETRIC ? p - 1 - m : p - m;
}
__device__ __forceinline__ int s2_remap(int style, int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    return s2_remap_slow(style, i, n);
}

// One input row r of a strip (u = r % RB, a literal after the caller's unrolling): stage 1 from the smem ring, stage 2
// through the register pipeline, emit the finished output row o = r - (Ly-1).  Only output rows 0 <= o < th are
// stored; everything else is computed and dropped.
// the multiply-accumulate of the kernel: common.cuh's mac (the reference's separate multiply and add for Float64), or — FMA,
// accum mode B2F_ACCUM_FMA — one fused multiply-add (half the FP64 instructions, one rounding instead of two)
template <typename CT, bool FMA> __device__ __forceinline__ CT s2_mac(CT acc, CT a, CT k) {
    if constexpr (FMA && sizeof(CT) == 8) return fma(a, k, acc);
    else return mac<CT>(acc, a, k);
}

template <typename CT, int LXT, int LYT, int LB, int NPL, int RB, int ROT, bool XS, bool YS, bool FMA>
__device__ __forceinline__ void s2_row(const int u, const int r, const S2Params<CT, NPL> &P, const int Lx, const int Ly,
                                       const int th, const CT *__restrict__ sblk, const int lane,
                                       const int tw, const bool lane_full, const bool lane_live,
                                       CT (&acc)[NPL][YS ? (LYT ? LYT : LB) : 1][S2Vec<CT>::PX],
                                       CT *(&outp)[NPL]) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int LBX = XS ? (LXT ? LXT : LB) : 1;
    constexpr int LBY = YS ? (LYT ? LYT : LB) : 1;
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX;
    constexpr int PW = 32 * PX + WIN;
    typedef typename S2Vec<CT>::T V;
    const CT *srow = sblk + (u % RB) * PW + lane * PX;
    CT v[WIN];
#pragma unroll
    for (int i = 0; i < WIN; i += PX) {
        if (LXT || i < PX + Lx - 1) {
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
    if constexpr (XS && std::is_same<CT, float>::value) {
        // Float32: input-major packed form.  Window value v[i] feeds outputs (2c, 2c+1) with the tap pair
        // (k[j], k[j-1]), j = i - 2c, as one FFMA2; the two end taps touch one output only (scalar FFMA).  Per output
        // the taps still arrive in ascending order, so the bits equal the scalar fmaf loop.
#pragma unroll
        for (int i = 0; i < PX + LBX - 1; ++i) {
#pragma unroll
            for (int c = 0; c < PX / 2; ++c) {
                const int j = i - 2 * c;
                if (j >= 0 && j <= LBX && (LXT || j <= Lx)) {
#pragma unroll
                    for (int p = 0; p < NPL; ++p) {
                        if (j == 0) {
                            mid[p][2 * c] = fmaf(v[i], P.kx[p][0], mid[p][2 * c]);
                        } else if (j < LBX && (LXT || j < Lx)) {
                            const float2 r = s2_fma2b(v[i], P.kxp[p][j], make_float2(mid[p][2 * c], mid[p][2 * c + 1]));
                            mid[p][2 * c] = r.x; mid[p][2 * c + 1] = r.y;
                        } else if (LXT || j == Lx) {
                            mid[p][2 * c + 1] = fmaf(v[i], P.kx[p][j - 1], mid[p][2 * c + 1]);
                        }
                    }
                }
            }
        }
    } else if constexpr (XS) {
#pragma unroll
        for (int j = 0; j < LBX; ++j) {
            if (LXT || j < Lx) {
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    const CT kj = P.kx[p][j];
#pragma unroll
                    for (int q = 0; q < PX; ++q) mid[p][q] = s2_mac<CT, FMA>(mid[p][q], v[q + j], kj);
                }
            }
        }
    } else {   // no stage along x: the row itself feeds stage 2
#pragma unroll
        for (int p = 0; p < NPL; ++p)
#pragma unroll
            for (int q = 0; q < PX; ++q) mid[p][q] = v[q];
    }
    // stage 2, transposed form (taps right-aligned in the LBY slots: a new output enters at slot LBY - Ly, whose
    // accumulator is never written and stays zero)
    CT fin[NPL][PX];
    if constexpr (YS) {
#pragma unroll
        for (int p = 0; p < NPL; ++p) {
            const CT kj = P.kyt[p][LBY - 1];
            if constexpr (std::is_same<CT, float>::value) {
#pragma unroll
                for (int q = 0; q < PX; q += 2) {
                    const float2 t = s2_fma2(make_float2(mid[p][q], mid[p][q + 1]), kj,
                                             make_float2(acc[p][LBY - 1][q], acc[p][LBY - 1][q + 1]));
                    fin[p][q] = t.x; fin[p][q + 1] = t.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < PX; ++q) fin[p][q] = s2_mac<CT, FMA>(acc[p][LBY - 1][q], mid[p][q], kj);
            }
        }
#pragma unroll
        for (int j = LBY - 2; j >= 0; --j) {
            if (LYT || j >= LBY - Ly) {
#pragma unroll
                for (int p = 0; p < NPL; ++p) {
                    const CT kj = P.kyt[p][j];
                    if constexpr (std::is_same<CT, float>::value) {
#pragma unroll
                        for (int q = 0; q < PX; q += 2) {
                            const float2 t = s2_fma2(make_float2(mid[p][q], mid[p][q + 1]), kj,
                                                     make_float2(acc[p][j][q], acc[p][j][q + 1]));
                            acc[p][j + 1][q] = t.x; acc[p][j + 1][q + 1] = t.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < PX; ++q) acc[p][j + 1][q] = s2_mac<CT, FMA>(acc[p][j][q], mid[p][q], kj);
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int p = 0; p < NPL; ++p)
#pragma unroll
            for (int q = 0; q < PX; ++q) fin[p][q] = mid[p][q];
    }
    // output row o = r-(Ly-1) is complete: emit it
    const int o = r - (Ly - 1);
    if (o >= 0 && o < th) {
        if (lane_full) {
#pragma unroll
            for (int p = 0; p < NPL; ++p) {
                V t;
#pragma unroll
                for (int q = 0; q < PX; ++q) ((CT *)&t)[q] = fin[p][q];
                *reinterpret_cast<V *>(outp[p]) = t;
            }
        } else if (lane_live) {
#pragma unroll
            for (int p = 0; p < NPL; ++p)
#pragma unroll
                for (int q = 0; q < PX; ++q)
                    if (lane * PX + q < tw) outp[p][q] = fin[p][q];
        }
#pragma unroll
        for (int p = 0; p < NPL; ++p) outp[p] += P.out_pitch;
    }
}

template <typename IT, typename CT, int LXT, int LYT, int LB, int NPL, int RB, int ROT, bool XS = true, bool YS = true, bool FMA = false>
__global__ void __launch_bounds__(S2_WARPS * 32) stream2d_kernel(const __grid_constant__ S2Params<CT, NPL> P) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int CW = 32 * PX;                              // strip width
    constexpr int LBX = XS ? (LXT ? LXT : LB) : 1;           // compile-time bound of the x taps (1: no x stage)
    constexpr int LBY = YS ? (LYT ? LYT : LB) : 1;
    constexpr int NCL = (CW + LBX - 1 + 31) / 32;            // loads per lane per input row
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX; // window registers (whole 128-bit granules)
    constexpr int PW = CW + WIN;                             // smem row pitch (elements), multiple of PX
    typedef typename S2Vec<CT>::T V;
    static_assert(LBX <= S2_MAXTAPS && LBY <= S2_MAXTAPS, "too many taps");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the strip geometry and the taps
    // in uniform registers and frees ~40 vector registers
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    CT *sbuf = reinterpret_cast<CT *>(smem_raw) + (size_t)warp * (2 * RB * PW);

    const long long sid = (long long)blockIdx.x * S2_WARPS + warp;
    if (sid >= P.nstrips) return;
    const int sx = (int)(sid % P.nsx);
    const int sy = (int)((sid / P.nsx) % P.nsy);
    const long long bz = sid / ((long long)P.nsx * P.nsy);

    const int Lx = XS ? (LXT ? LXT : P.Lx) : 1;
    const int Ly = YS ? (LYT ? LYT : P.Ly) : 1;
    const int x0 = P.rx0 + sx * CW;
    const int y0 = P.ry0 + sy * P.SH;
    const int tw = min(CW, P.rx0 + P.rw - x0);          // live output columns
    const int th = min(P.SH, P.ry0 + P.rh - y0);        // live output rows
    const int in_rows = th + Ly - 1;
    const int in_cols = CW + Lx - 1;
    const IT *__restrict__ img = reinterpret_cast<const IT *>(P.img) + bz * P.img_plane;
    const bool is_fill = P.style == B2F_FILL;

    // Source column of each of this lane's load slots, remapped through the border (src/border.jl:564-590).
    // Loads are unconditional (slots without a pixel read column 0) and patched when parked.
    int gx[NCL];
    unsigned colfill = 0, coldead = 0;
#pragma unroll
    for (int c = 0; c < NCL; ++c) {
        const int col = lane + 32 * c;
        int g = 0;
        if (col < in_cols) {
            g = s2_remap(P.style, x0 + (XS ? P.klox : 0) + col, P.W);
            if (g < 0) { colfill |= 1u << c; g = 0; }
        } else {
            coldead |= 1u << c;
        }
        gx[c] = g;
    }
    const int ytop = y0 + (YS ? P.kloy : 0);              // buffer row of strip-local input row 0
    const bool y_interior = ytop >= 0 && ytop + in_rows <= P.H;

    IT stage[RB][NCL];
    unsigned rowfill = 0;       // bit rr: staged row rr lies in the Fill region
    auto fetch_block = [&](int blk) {
        rowfill = 0;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
            const int r = min(blk * RB + rr, in_rows - 1);   // rows past the strip are never used: clamp the address
            int gy = ytop + r;
            if (!y_interior && (unsigned)gy >= (unsigned)P.H) {
                // outside this buffer: apply the border in GLOBAL row coordinates (slab form; y_first = 0, Hg = H otherwise)
                gy = s2_remap(P.style, gy + P.y_first, P.Hg);
                if (gy < 0) { rowfill |= 1u << rr; gy = 0; } else gy -= P.y_first;
            }
            const IT *row = img + (long long)gy * P.W;
#pragma unroll
            for (int c = 0; c < NCL; ++c) stage[rr][c] = __ldg(row + gx[c]);
        }
    };
    auto park_block = [&](int blk) {
        CT *dst = sbuf + (blk & 1) * (RB * PW) + lane;
#pragma unroll
        for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
            for (int c = 0; c < NCL; ++c) {
                CT v = S2Conv<IT, CT>::f(stage[rr][c], P.n0_r, P.n0_c);
                if (is_fill && ((colfill >> c | rowfill >> rr) & 1u)) v = P.fill;
                if (c < NCL - 1 || !((coldead >> c) & 1u)) dst[rr * PW + 32 * c] = v;
            }
        }
    };

    CT acc[NPL][LBY][PX];       // acc[.][0] stays zero: the entry slot of a new output
#pragma unroll
    for (int p = 0; p < NPL; ++p)
#pragma unroll
        for (int s = 0; s < LBY; ++s)
#pragma unroll
            for (int q = 0; q < PX; ++q) acc[p][s][q] = (CT)0;

    // output pointers of the row being emitted next (row o = 0 first)
    CT *outp[NPL];
#pragma unroll
    for (int p = 0; p < NPL; ++p)
        outp[p] = reinterpret_cast<CT *>(P.out[p]) + bz * P.out_plane + (long long)(y0 - P.out_oy) * P.out_pitch +
                  (x0 + lane * PX - P.out_ox);
    const bool lane_full = P.vec_ok && (lane * PX + PX <= tw);
    const bool lane_live = lane * PX < tw;

    fetch_block(0);
    park_block(0);
    __syncwarp();

    // ---- main loop: one prefetch block of RB rows per iteration, no per-row checks: rows past the end of the strip are
    // clamped duplicates whose contributions only reach output rows that are never stored (o >= th)
    for (int rbase = 0, blk = 0; rbase < in_rows; rbase += RB, ++blk) {
        fetch_block(blk + 1);                                     // loads fly while this block is computed
#pragma unroll
        for (int u = 0; u < RB; ++u)
            s2_row<CT, LXT, LYT, LB, NPL, RB, ROT, XS, YS, FMA>(u, rbase + u, P, Lx, Ly, th, sbuf + (blk & 1) * (RB * PW), lane, tw, lane_full,
                                                           lane_live, acc, outp);
        park_block(blk + 1);
        __syncwarp();
    }
}

}  // namespace b2f
