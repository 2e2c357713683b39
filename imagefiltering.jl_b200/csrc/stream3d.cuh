// stream3d.cuh — K1-3D: fused 3-D separable FIR (x, y, z stages in ONE launch), Float32, slab-aware.
//
// Replaces, for a cascade of three 1-D factors on axes 1,2,3 (KernelFactors.gaussian((s,s,s)) & friends), the
// reference's padarray + three full-volume passes through padded temporaries (src/imfilter.jl:321-341, 385-395,
// 438-446, loops :724-739): every input voxel is read from HBM once, every output voxel written once.
//
// Config 5 (17+17+17 taps, 8 B/voxel) sits on the FP32 ridge of the machine, and the binding on-chip resources are the
// FMA issue slots and the shared-memory bandwidth, so the design minimises both per voxel:
//
//   * a CTA (512 threads, all warps alike — no role specialisation, one split mbarrier hand-off per plane) owns a
//     32 x 64 tile of the xy-plane and MARCHES along z over its chunk of planes;
//   * raw planes (tile + halo, 48 x 80 floats) arrive through an 8-deep TMA ring (cp.async.bulk.tensor, mbarrier
//     completion, issued 8 planes ahead by one thread; out-of-range cells read as zero).  Border tiles are patched in
//     shared memory one plane before use from a per-CTA gather list built once through the border remap
//     (src/border.jl:564-590 semantics in x, y; z is remapped by choosing the source plane): "pad the input once",
//     exactly the reference's semantics, with no padded copy;
//   * stage x: each thread makes 8 adjacent outputs of one raw row from a 24-value register window (6 LDS.128,
//     conflict-free because the tile pitches are an odd number of 16-byte chunks) -> xf tile (double buffered);
//   * stage y: each thread owns 2 columns x 2 rows: 18 LDS.64 feed 2 float2 outputs from a register window;
//   * stage z runs in TRANSPOSED (systolic) form, entirely in registers: the thread keeps the Lz-1 partial sums of each
//     of its 4 voxel columns (2 x 16 float2); a new xy-filtered value m completes the oldest output (stored at once, one
//     8-byte store per row) and every other partial sum moves one slot up while it takes its tap,
//     acc[j+1] = m * k[j] + acc[j].  No z ring in shared memory, no z re-reads, no register moves, static indices
//     without unrolling the plane loop (an Lz-fold unrolled loop overflows the 32 KB instruction cache: measured 2
//     no_instruction stalls per issue);
//   * every tap loop is unrolled with static register indices and ascending tap order per output (the reference's
//     order); all multiply-adds are packed FFMA2 (fma.rn.f32x2, tap broadcast from a uniform register).
//
// Slab form (multi-GPU): the planes of the last axis may live in three buffers — this rank's owned planes
// plus `lo`/`hi` halo planes, which are either receive buffers or PEER memory of the neighbouring GPU mapped over
// NVLink (the kernel then performs the halo exchange itself, by TMA / P2P loads, fused with the filter).
#pragma once

#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace b2f {

constexpr int S3_TX = 32, S3_TY = 64;   // tile (outputs)
constexpr int S3_R = 2;                 // rows per thread in stages y / z
constexpr int S3_MAXTAPS = 17;
constexpr int S3_RWP = 52;              // raw tile pitch in floats: 13 chunks of 16 B (odd)
constexpr int S3_XFP = 36;              // x-filtered tile pitch: 9 chunks (odd)
constexpr int S3_NT = 512;
constexpr int S3_NRAW = 8;              // raw ring depth (power of two)
constexpr int S3_NXF = 3;               // x-filtered tile ring depth
constexpr int S3_AHEAD = S3_NRAW - 2;   // the TMA of plane p is issued in interval p - 2 - AHEAD
constexpr int S3_PT = 64, S3_PTA = 32, S3_PTB = 16;   // plane-source ring: entries, look-ahead and block of its refill

struct S3Params {
    const float *own, *lo, *hi;    // owned planes / halo planes below / above (dense W x H planes)
    int own_first, own_n, lo_n, hi_n, Zg;   // global index of the first owned plane, plane counts, global extent
    int W, H;
    long long plane;               // W * H
    float *out;                    // owned planes only
    int style;
    float fill;
    int Lx, Ly, Lz, klox, kloy, kloz;
    int zchunk, ntx, nty;          // output planes per z-chunk, tiles along x / y
    int nfull, kch;                // the first nfull tiles march all planes in one CTA, the others are cut into kch chunks
    int vec_out;                   // 8-byte stores are aligned
    int xsh;                       // tiles start at x = 32*tx - xsh, so that the TMA box starts on a 16-byte boundary
    // staged halos: lo / hi are LOCAL buffers being filled by copy engines while this kernel runs; *flag == epoch once the
    // buffer is complete (NULL: the planes are there already)
    const unsigned char *flag_lo, *flag_hi;   // flag_lo[0]: rows [0, lo_early_rows) of every lower halo plane, flag_lo[1]: the rest
    int epoch, lo_early_rows;
    int use_tma;                   // the tensor maps are valid (else every cell comes through the gather loader)
    float kx[S3_MAXTAPS], ky[S3_MAXTAPS];
    float kzr[S3_MAXTAPS];         // z taps RIGHT-aligned in the instantiation's LBZ slots
    float2 kxp[S3_MAXTAPS];        // kxp[j] = (kx[j], kx[j-1]): the taps one input value carries to two adjacent outputs
    float2 kyp[S3_MAXTAPS];        // the same pairs of the y taps (v2: a thread owns 4 consecutive rows of one column)
};

__device__ __forceinline__ float2 s3_fma2(float2 a, float k, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rc = *reinterpret_cast<unsigned long long *>(&c), rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(k));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
// (v*k.x + c.x, v*k.y + c.y): one value, two taps (SASS: FFMA2 with the .F32 broadcast form of operand a)
__device__ __forceinline__ float2 s3_fma2b(float v, float2 k, float2 c) {
    unsigned long long rk = *reinterpret_cast<unsigned long long *>(&k), rc = *reinterpret_cast<unsigned long long *>(&c), rv, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rv), "l"(rk), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// source of global plane index zi: halo buffers are matched on the LOGICAL index first (a circular wrap arrives as
// an ordinary halo), then the border remap is applied in global coordinates.  which: 0 own, 1 lo, 2 hi, -1 Fill plane.
__device__ __forceinline__ void s3_locate(const S3Params &P, int zi, int &which, int &zc) {
    int rel = zi - P.own_first;
    which = 0;
    if ((unsigned)rel < (unsigned)P.own_n) { zc = rel; return; }
    if (rel < 0 && rel >= -P.lo_n) { which = 1; zc = rel + P.lo_n; return; }
    if (rel >= P.own_n && rel < P.own_n + P.hi_n) { which = 2; zc = rel - P.own_n; return; }
    const int g = (int)remap_index(P.style, (int64_t)zi, (int64_t)P.Zg);
    rel = g - P.own_first;
    if (g < 0) { which = -1; zc = 0; }
    else if ((unsigned)rel < (unsigned)P.own_n) { zc = rel; }
    else if (rel < 0) { which = 1; zc = rel + P.lo_n; }        // host validated that it is present
    else { which = 2; zc = rel - P.own_n; }
}

// ---- mbarrier / TMA primitives (sm_90+ PTX) ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned s3_sa(const void *p) {
    // through an opaque move: the compiler otherwise REMATERIALISES the address (S2R CgaCtaId + LEA) at every use
    unsigned a;
    asm volatile("mov.u32 %0, %1;" : "=r"(a) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return a;
}
// All of these take 32-bit shared-window addresses computed ONCE per CTA (s3_sa): converting a generic pointer costs an
// S2R of the CTA-in-cluster id every time, seven special-register reads per plane in the first version of the loop.
__device__ __forceinline__ void s3_mbar_init(unsigned b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory");
}
__device__ __forceinline__ void s3_mbar_arrive(unsigned b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void s3_mbar_expect_tx(unsigned b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint expires) instead of
// having it re-poll — in the first version the polling loop was 6 % of all issued instructions
__device__ __forceinline__ void s3_mbar_wait(unsigned b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(b), "r"(parity), "r"(200000u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void s3_tma_load3d(unsigned dst, const void *tmap, unsigned bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// shared-memory accesses of the hot stages by 32-bit shared address (same reason as above: no generic-pointer conversions)
__device__ __forceinline__ float4 s3_lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 s3_lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float s3_lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void s3_sts128(unsigned a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// compile-time geometry
template <int LXT, int LYT, int LZT> struct S3C {
    static constexpr int LBX = LXT ? LXT : S3_MAXTAPS, LBY = LYT ? LYT : S3_MAXTAPS, LBZ = LZT ? LZT : S3_MAXTAPS;
    static constexpr int RH = S3_TY + LBY - 1;                    // raw / xf tile rows (compile-time bound), <= 64
    static constexpr int RAWBYTES = RH * S3_RWP * 4;              // one TMA box
    static constexpr int RAWSZ = ((RAWBYTES + 127) / 128) * 32;   // buffer stride in floats (128-byte multiple)
    static constexpr int XFSZ = RH * S3_XFP;
    static constexpr int WINX = ((8 + LBX - 1 + 3) / 4) * 4;      // x window registers (whole 16-byte chunks), <= 24
    static constexpr int NCELL = ((RH * (S3_TX + LBX - 1) + 3) / 4) * 4;   // gather list capacity: the whole raw tile
    static constexpr size_t SMEM = sizeof(float) * (size_t)(S3_NRAW * RAWSZ + S3_NXF * XFSZ) + (sizeof(int) + sizeof(short)) * NCELL +
                                   sizeof(int) * 2 * S3_PT + sizeof(uint64_t) * (S3_NRAW + S3_NXF);
    static_assert(RH <= 8 * (S3_NT / 32) && S3_TX + LBX - 1 <= S3_RWP && 8 * 3 + WINX <= S3_RWP, "tile geometry");
};

// ---- stage x: one row group of 8 outputs per thread.  A quarter-warp covers two rows x 32 columns ----------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_x_task(const S3Params &P, const unsigned rb, const unsigned xb,
                                          const int in_rows, const int Lx, const int warp, const int lane) {
    typedef S3C<LXT, LYT, LZT> C;
    const int l8 = lane & 7, qw = lane >> 3;
    const int xg = l8 & 3, row = 8 * warp + 2 * qw + (l8 >> 2);
    if (row >= in_rows) return;
    const unsigned src = rb + (unsigned)(row * S3_RWP + 8 * xg) * 4u;
    float v[C::WINX];
#pragma unroll
    for (int i = 0; i < C::WINX; i += 4) {
        const float4 t = s3_lds128(src + i * 4);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
    // input-major: window value v[i] feeds outputs (2c, 2c+1) with the tap pair (k[j], k[j-1]), j = i - 2c (one FFMA2
    // with the value broadcast); the two end taps touch one output only (scalar FFMA).  Per output the taps still
    // arrive in ascending order.
    float2 a[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8 + C::LBX - 1; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = i - 2 * c;
            if (j >= 0 && j <= C::LBX && (LXT || j <= Lx)) {
                if (j == 0) a[c].x = fmaf(v[i], P.kx[0], a[c].x);
                else if (j < C::LBX && (LXT || j < Lx)) a[c] = s3_fma2b(v[i], P.kxp[j], a[c]);
                else if (LXT || j == Lx) a[c].y = fmaf(v[i], P.kx[j - 1], a[c].y);
            }
        }
    }
    const unsigned d = xb + (unsigned)(row * S3_XFP + 8 * xg) * 4u;
    s3_sts128(d, make_float4(a[0].x, a[0].y, a[1].x, a[1].y));
    s3_sts128(d + 16, make_float4(a[2].x, a[2].y, a[3].x, a[3].y));
}

// ---- stage y: 2 columns x S3_R rows per thread; a half-warp covers the 32 columns of one row group -----------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_y_task(const S3Params &P, const unsigned xb, float2 (&m)[S3_R], const int Ly) {
    typedef S3C<LXT, LYT, LZT> C;
#pragma unroll
    for (int o = 0; o < S3_R; ++o) m[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < S3_R + C::LBY - 1; ++i) {
        if (LYT || i < S3_R + Ly - 1) {
            const float2 s = s3_lds64(xb + i * (S3_XFP * 4));
#pragma unroll
            for (int o = 0; o < S3_R; ++o) {
                const int j = i - o;
                if (j >= 0 && j < C::LBY && (LYT || j < Ly)) m[o] = s3_fma2(s, P.ky[j], m[o]);
            }
        }
    }
}

// ---- stage z in transposed (systolic) form: acc[j] is the partial sum of the output that takes tap j next.  The new
// xy-filtered value m completes the oldest output (tap LBZ-1: stored) and moves every other partial sum one slot up
// WHILE adding its tap — acc[j+1] = m * k[j] + acc[j] — so the rotation costs nothing: all register indices are static,
// the plane loop is not unrolled, and each output still receives its taps in ascending order.  Run-time tap counts are
// right-aligned in the LBZ slots: a new output enters at slot LBZ - Lz (the slots below stay zero) ---------------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_z_update(const S3Params &P, float2 (&acc)[S3_R][S3C<LXT, LYT, LZT>::LBZ],
                                            const float2 (&m)[S3_R], const int Lz, float *__restrict__ op, const int W,
                                            const int nrow, const int smode, const bool emit) {
    typedef S3C<LXT, LYT, LZT> C;
    float2 fin[S3_R];
    {
        const float k = P.kzr[C::LBZ - 1];
#pragma unroll
        for (int o = 0; o < S3_R; ++o) fin[o] = s3_fma2(m[o], k, acc[o][C::LBZ - 1]);
    }
#pragma unroll
    for (int j = C::LBZ - 2; j >= 0; --j) {
        if (LZT || j >= C::LBZ - Lz) {
            const float k = P.kzr[j];
#pragma unroll
            for (int o = 0; o < S3_R; ++o) acc[o][j + 1] = s3_fma2(m[o], k, acc[o][j]);
        }
    }
    if (emit) {
        if (smode == 3) {
#pragma unroll
            for (int o = 0; o < S3_R; ++o)
                if (o < nrow) *reinterpret_cast<float2 *>(op + o * W) = fin[o];
        } else if (smode != 0) {
#pragma unroll
            for (int o = 0; o < S3_R; ++o) {
                if (o < nrow) {
                    if (smode != 4) op[o * W] = fin[o].x;
                    if (smode != 1) op[o * W + 1] = fin[o].y;
                }
            }
        }
    }
}

// LXT/LYT/LZT > 0: exact tap counts; 0: run-time count bounded by S3_MAXTAPS (uniform predicates)
template <int LXT, int LYT, int LZT>
__global__ void __launch_bounds__(S3_NT, 1)
stream3d_kernel(const __grid_constant__ S3Params P, const __grid_constant__ CUtensorMap m_own,
                const __grid_constant__ CUtensorMap m_lo, const __grid_constant__ CUtensorMap m_hi) {
    typedef S3C<LXT, LYT, LZT> C;
    constexpr int TX = S3_TX, TY = S3_TY, RAWSZ = C::RAWSZ, XFSZ = C::XFSZ, LBZ = C::LBZ, N = S3_NRAW;

    extern __shared__ __align__(1024) float s3_smem[];
    float *raw = s3_smem;                       // N x RAWSZ
    float *xf = raw + N * RAWSZ;                // S3_NXF x XFSZ
    int *cell_src = reinterpret_cast<int *>(xf + S3_NXF * XFSZ);          // gather list: source offset inside a plane (-1: Fill)
    unsigned short *cell_dst = reinterpret_cast<unsigned short *>(cell_src + C::NCELL);   // ... and raw-tile offset
    int *ptw = reinterpret_cast<int *>(cell_dst + C::NCELL);         // ring of plane sources: buffer (0 own, 1 lo, 2 hi, -1 Fill)
    int *ptz = ptw + S3_PT;                                           // ... and plane index inside it
    // barriers, as 32-bit shared addresses: full + 8 b = TMA of raw buffer b landed; xfull + 8 s = every warp is through
    // stage x of the plane in xf slot s
    const unsigned full = s3_sa(ptz + S3_PT), xfull = full + 8 * N, raw_sa = s3_sa(raw), xf_sa = s3_sa(xf);

    // the thread index through an opaque move: the compiler otherwise re-reads the special register (S2R, ~20 cycles on the
    // critical path) four times per plane instead of keeping it in a register
    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5, lane = tid & 31;
    const int Lx = LXT ? LXT : P.Lx, Ly = LYT ? LYT : P.Ly, Lz = LZT ? LZT : P.Lz;
    const int bid = blockIdx.x;
    int tile = bid, ch = 0, zc = P.own_n;
    if (bid >= P.nfull) {
        const int b2 = bid - P.nfull;
        tile = P.nfull + b2 / P.kch;
        ch = b2 - (tile - P.nfull) * P.kch;
        zc = P.zchunk;
    }
    const int tx = tile % P.ntx, ty = tile / P.ntx;
    const int x0 = tx * TX - P.xsh, y0 = ty * TY;
    const int in_cols = TX + Lx - 1, in_rows = TY + Ly - 1;
    const int zo0 = P.own_first + ch * zc;                               // first output plane of this chunk (global)
    const int nout = min(zc, P.own_first + P.own_n - zo0);
    if (nout <= 0) return;
    const int in_planes = nout + Lz - 1;
    const int zin0 = zo0 + P.kloz;                                       // global index of input plane p = 0
    const int xa = x0 + P.klox, ya = y0 + P.kloy;
    const bool tma = P.use_tma != 0;

    // in-range part of the raw tile: columns [cl, cr), rows [rt, rb); every other cell goes on the gather list
    int cl = min(max(-xa, 0), in_cols), cr = min(max(P.W - xa, 0), in_cols);
    const int rt = tma ? min(max(-ya, 0), in_rows) : 0, rb = tma ? min(max(P.H - ya, 0), in_rows) : in_rows;
    if (!tma) cl = cr = in_cols;                                         // every column is "out of range": gather it all
    const bool fix = !tma || (P.style != B2F_FILL && (cl > 0 || cr < in_cols || rt > 0 || rb < in_rows));
    // column strips [0,cl) u [cr,in_cols) on every row, then row strips [0,rt) u [rb,in_rows) on the in-range columns
    const int ncs = cl + (in_cols - cr), n1 = ncs * in_rows, wc = cr - cl, ncell = fix ? n1 + (rt + (in_rows - rb)) * wc : 0;
    for (int idx = tid; idx < ncell; idx += S3_NT) {
        int r, c;
        if (idx < n1) {
            r = idx / ncs;
            const int k = idx - r * ncs;
            c = k < cl ? k : cr + (k - cl);
        } else {
            const int i2 = idx - n1;
            const int rr = i2 / wc;
            c = cl + (i2 - rr * wc);
            r = rr < rt ? rr : rb + (rr - rt);
        }
        const int sx = (int)remap_index(P.style, (int64_t)xa + c, (int64_t)P.W);
        const int sy = (int)remap_index(P.style, (int64_t)ya + r, (int64_t)P.H);
        cell_src[idx] = (sx < 0 || sy < 0) ? -1 : sy * P.W + sx;
        cell_dst[idx] = (unsigned short)(r * S3_RWP + c);
    }
    if (tid == 0) {
        for (int i = 0; i < N; ++i) s3_mbar_init(full + 8 * i, 1);
        for (int i = 0; i < S3_NXF; ++i) s3_mbar_init(xfull + 8 * i, S3_NT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const int qb = -2;                          // first interval of the plane loop
    // plane sources: entry p & (S3_PT-1) describes input plane p; filled S3_PTA planes ahead of the loop
    auto locate_block = [&](int p0, int n) {
        if (tid < n) {
            const int p = p0 + tid;
            int which = -1, zc = 0;
            if (p >= 0 && p < in_planes) s3_locate(P, zin0 + p, which, zc);
            ptw[p & (S3_PT - 1)] = which;
            ptz[p & (S3_PT - 1)] = zc;
        }
    };
    locate_block(qb, S3_PTA);               // planes -2 .. PTA-3; the loop refills PTB planes at a time, PTA ahead
    __syncthreads();

    bool lo_ready = P.flag_lo == nullptr, hi_ready = P.flag_hi == nullptr;     // thread 0 only
    auto wait_flag = [&](const unsigned char *f) {
        // bounded: a copy that never arrives (a failed transfer on the side stream) must become an error, not a hung GPU
        unsigned long long t0 = 0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*reinterpret_cast<const volatile unsigned char *>(f) != (unsigned char)P.epoch) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 5000000000ULL) asm volatile("trap;");          // 5 s
        }
        __threadfence_system();
    };
    auto issue = [&](int p) {                   // one thread: TMA of input plane p into its ring buffer
        const int which = ptw[p & (S3_PT - 1)];
        if (which == 1 && !lo_ready) { wait_flag(P.flag_lo + (ya + in_rows > P.lo_early_rows ? 1 : 0)); lo_ready = true; }
        if (which == 2 && !hi_ready) { wait_flag(P.flag_hi); hi_ready = true; }
        const int zc = which < 0 ? P.own_n : ptz[p & (S3_PT - 1)];         // Fill(0) plane: out of range reads zero
        const void *map = which == 1 ? (const void *)&m_lo : which == 2 ? (const void *)&m_hi : (const void *)&m_own;
        const int b = p & (N - 1);
        s3_mbar_expect_tx(full + 8 * b, (unsigned)C::RAWBYTES);
        s3_tma_load3d(raw_sa + b * (RAWSZ * 4), map, full + 8 * b, xa, ya, zc);
    };
    auto fixup = [&](int p) {                   // all threads: the gather list of input plane p
        const int which = ptw[p & (S3_PT - 1)];
        const float *src = which < 0 ? nullptr
                                     : (which == 1 ? P.lo : which == 2 ? P.hi : P.own) + (long long)ptz[p & (S3_PT - 1)] * P.plane;
        float *dst = raw + (p & (N - 1)) * RAWSZ;
        for (int base = tid; base < ncell; base += 4 * S3_NT) {       // four gathers in flight per thread
            float v[4];
            int d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = base + k * S3_NT;
                d[k] = -1;
                if (idx < ncell) {
                    const int so = cell_src[idx];
                    d[k] = cell_dst[idx];
                    v[k] = (src != nullptr && so >= 0) ? __ldg(src + so) : P.fill;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (d[k] >= 0) dst[d[k]] = v[k];
        }
        if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };

    if (tma && tid == 0)
        for (int p = 0; p < min(S3_AHEAD, in_planes); ++p) issue(p);

    // this thread's 2 x S3_R voxel columns in stages y / z
    const int pc = tid & 15, rg = tid >> 4;
    const int gx = x0 + 2 * pc, gy = y0 + S3_R * rg;
    // 3: 8-byte store, 2: two scalars, 1: the first only, 4: the second only, 0: none
    const int smode = (gx >= P.W || gx < -1) ? 0 : gx == -1 ? 4 : (gx + 1 >= P.W ? 1 : (P.vec_out ? 3 : 2));
    const int nrow = min(S3_R, P.H - gy);                                           // <= 0: nothing to store
    // output pointer of the plane completed by input plane q: advanced by one plane per interval
    float *op = P.out + ((long long)(zo0 - P.own_first) + (qb - (Lz - 1))) * P.plane + (long long)gy * P.W + gx;
    const int yoff = (S3_R * rg) * S3_XFP + 2 * pc;

    float2 acc[S3_R][LBZ];
#pragma unroll
    for (int o = 0; o < S3_R; ++o)
#pragma unroll
        for (int j = 0; j < LBZ; ++j) acc[o][j] = make_float2(0.f, 0.f);

    // Interval q: gather-patch plane q+2, stage x on plane q+1, stages y+z on plane q; TMA of plane q+2+AHEAD goes out.
    // The only CTA-wide synchronisation is a SPLIT barrier per plane: a warp arrives on xfull[p % 3] when it is through
    // stage x of plane p (and, in program order before that, the gather of plane p+1 and stages y/z of plane p-2), and
    // waits for it one interval later, just before it needs plane p in stage y.  That one hand-off orders everything:
    // xf[p % 3] is complete, xf[(p+1) % 3] (plane p-2) and raw[p % 8] (plane p) are free, the gather of p+1 has landed.
    // Between arrive and wait lies a whole y/z stage, so warps drift apart by up to a plane instead of idling in lockstep.
    // Stage x needs XW of the NW warps; even planes take warps 0.., odd planes warps NW-XW.., which evens the FMA load
    // of the four schedulers over two planes.
    constexpr int NW = S3_NT / 32, XW = (C::RH + 7) / 8, XOFF = NW > XW ? NW - XW : 0;
    int wb = 0, wph = 0, ab = 0;                // xf slot / phase of plane q, xf slot of plane q + 1
    for (int q = -2; q < in_planes; ++q) {
        if (((q + 2) & (S3_PTB - 1)) == 0) locate_block(q + S3_PTA, S3_PTB);   // read >= S3_PTA - S3_AHEAD - 2 intervals later
        if (q >= 0) s3_mbar_wait(xfull + 8 * wb, wph);
        if (tma && tid == 0 && q + 2 + S3_AHEAD < in_planes) issue(q + 2 + S3_AHEAD);
        if (fix && q + 2 < in_planes) {
            const int p = q + 2;
            if (tma) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
            fixup(p);
        }
        if (q + 1 < in_planes && q + 1 >= 0) {
            const int p = q + 1;
            const int xw = (p & 1) ? warp - XOFF : warp;
            if (xw >= 0) {
                if (tma && !fix) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
                s3_x_task<LXT, LYT, LZT>(P, raw_sa + (p & (N - 1)) * (RAWSZ * 4), xf_sa + ab * (XFSZ * 4), in_rows, Lx, xw, lane);
            }
            __syncwarp();
            if (lane == 0) s3_mbar_arrive(xfull + 8 * ab);
            ab = ab == S3_NXF - 1 ? 0 : ab + 1;
        }
        if (q < 0) {
            __syncthreads();                    // prologue: the gathers of planes 0 and 1 land before stage x reads them
        } else {
            float2 m[S3_R];
            s3_y_task<LXT, LYT, LZT>(P, xf_sa + (wb * XFSZ + yoff) * 4, m, Ly);
            s3_z_update<LXT, LYT, LZT>(P, acc, m, Lz, op, P.W, nrow, smode, q >= Lz - 1 && nrow > 0);
            if (wb == S3_NXF - 1) { wb = 0; wph ^= 1; } else ++wb;
        }
        op += P.plane;
    }
}

// =====================================================================================================================
// v2 of the marching kernel (round 2).  Same tile, rings, barriers and stage x as above; what changed:
//   * stages y / z: a thread owns ONE column x FOUR consecutive rows (warp w = rows 4w..4w+3, lane = column).  Stage y
//     reads 4 + Ly - 1 single floats of its column (LDS.32, one wavefront per warp and row: 20 B per voxel instead of
//     the 36 B of the 2 x 2 mapping — shared-memory bandwidth was the tightest floor of v1) and feeds two row pairs in
//     the value-broadcast x tap-pair form of stage x; stage z keeps the pairs (rows 0,1) and (rows 2,3) as float2
//     partial sums; a finished plane leaves as four 128-byte row segments per warp;
//   * the plane loop is split into ramp-up, STEADY STATE and drain: the steady-state body has no per-plane predicates
//     (every sub-step is active, border patching is a compile-time flag), which removes the ~25 branches, the ISETPs and
//     the BSSY/BSYNC pairs v1 executed per warp and plane.
// =====================================================================================================================
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_y_task4(const S3Params &P, const unsigned xb, float2 (&m)[2], const int Ly) {
    typedef S3C<LXT, LYT, LZT> C;
    m[0] = make_float2(0.f, 0.f);
    m[1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4 + C::LBY - 1; ++i) {
        if (LYT || i < 4 + Ly - 1) {
            const float s = s3_lds32(xb + i * (S3_XFP * 4));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = i - 2 * h;            // row pair h = outputs (2h, 2h+1): taps (j, j-1) of input row i
                if (j >= 0 && j <= C::LBY && (LYT || j <= Ly)) {
                    if (j == 0) m[h].x = fmaf(s, P.ky[0], m[h].x);
                    else if (j < C::LBY && (LYT || j < Ly)) m[h] = s3_fma2b(s, P.kyp[j], m[h]);
                    else if (LYT || j == Ly) m[h].y = fmaf(s, P.ky[j - 1], m[h].y);
                }
            }
        }
    }
}

template <int LXT, int LYT, int LZT, bool CS>
__device__ __forceinline__ void s3_z_update4(const S3Params &P, float2 (&acc)[2][S3C<LXT, LYT, LZT>::LBZ], const float2 (&m)[2],
                                             const int Lz, float *__restrict__ op, const int W, const int nrow, const bool emit) {
    typedef S3C<LXT, LYT, LZT> C;
    float2 fin[2];
    // the two 64-bit register operands of an FFMA2 must come from different register banks (bank = bit 1 of the register
    // number); consecutive partial sums alternate banks, so the new value is kept in both: without the explicit second copy
    // ptxas re-copies it for every other tap (~16 MOVs per plane in v1)
    float2 mb[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        asm volatile("mov.b32 %0, %1;" : "=f"(mb[o].x) : "f"(m[o].x));
        asm volatile("mov.b32 %0, %1;" : "=f"(mb[o].y) : "f"(m[o].y));
    }
    {
        const float k = P.kzr[C::LBZ - 1];
#pragma unroll
        for (int o = 0; o < 2; ++o) fin[o] = s3_fma2(((C::LBZ - 1) & 1) ? mb[o] : m[o], k, acc[o][C::LBZ - 1]);
    }
#pragma unroll
    for (int j = C::LBZ - 2; j >= 0; --j) {
        if (LZT || j >= C::LBZ - Lz) {
            const float k = P.kzr[j];
#pragma unroll
            for (int o = 0; o < 2; ++o) acc[o][j + 1] = s3_fma2((j & 1) ? mb[o] : m[o], k, acc[o][j]);
        }
    }
    if (emit) {
        const float v[4] = {fin[0].x, fin[0].y, fin[1].x, fin[1].y};
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (o < nrow) {
                if (CS) __stcs(op + (long long)o * W, v[o]); else op[(long long)o * W] = v[o];
            }
    }
}

template <int LXT, int LYT, int LZT, bool CS>
__global__ void __launch_bounds__(S3_NT, 1)
stream3d_kernel2(const __grid_constant__ S3Params P, const __grid_constant__ CUtensorMap m_own,
                 const __grid_constant__ CUtensorMap m_lo, const __grid_constant__ CUtensorMap m_hi) {
    typedef S3C<LXT, LYT, LZT> C;
    constexpr int TX = S3_TX, TY = S3_TY, RAWSZ = C::RAWSZ, XFSZ = C::XFSZ, LBZ = C::LBZ, N = S3_NRAW;

    extern __shared__ __align__(1024) float s3_smem[];
    float *raw = s3_smem;                       // N x RAWSZ
    float *xf = raw + N * RAWSZ;                // S3_NXF x XFSZ
    int *cell_src = reinterpret_cast<int *>(xf + S3_NXF * XFSZ);
    unsigned short *cell_dst = reinterpret_cast<unsigned short *>(cell_src + C::NCELL);
    int *ptw = reinterpret_cast<int *>(cell_dst + C::NCELL);
    int *ptz = ptw + S3_PT;
    const unsigned full = s3_sa(ptz + S3_PT), xfull = full + 8 * N, raw_sa = s3_sa(raw), xf_sa = s3_sa(xf);

    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5, lane = tid & 31;
    const int Lx = LXT ? LXT : P.Lx, Ly = LYT ? LYT : P.Ly, Lz = LZT ? LZT : P.Lz;
    const int bid = blockIdx.x;
    int tile = bid, ch = 0, zc = P.own_n;
    if (bid >= P.nfull) {
        const int b2 = bid - P.nfull;
        tile = P.nfull + b2 / P.kch;
        ch = b2 - (tile - P.nfull) * P.kch;
        zc = P.zchunk;
    }
    const int tx = tile % P.ntx, ty = tile / P.ntx;
    const int x0 = tx * TX - P.xsh, y0 = ty * TY;
    const int in_cols = TX + Lx - 1, in_rows = TY + Ly - 1;
    const int zo0 = P.own_first + ch * zc;
    const int nout = min(zc, P.own_first + P.own_n - zo0);
    if (nout <= 0) return;
    const int in_planes = nout + Lz - 1;
    const int zin0 = zo0 + P.kloz;
    const int xa = x0 + P.klox, ya = y0 + P.kloy;
    const bool tma = P.use_tma != 0;

    int cl = min(max(-xa, 0), in_cols), cr = min(max(P.W - xa, 0), in_cols);
    const int rt = tma ? min(max(-ya, 0), in_rows) : 0, rb = tma ? min(max(P.H - ya, 0), in_rows) : in_rows;
    if (!tma) cl = cr = in_cols;
    const bool fix = !tma || (P.style != B2F_FILL && (cl > 0 || cr < in_cols || rt > 0 || rb < in_rows));
    const int ncs = cl + (in_cols - cr), n1 = ncs * in_rows, wc = cr - cl, ncell = fix ? n1 + (rt + (in_rows - rb)) * wc : 0;
    for (int idx = tid; idx < ncell; idx += S3_NT) {
        int r, c;
        if (idx < n1) {
            r = idx / ncs;
            const int k = idx - r * ncs;
            c = k < cl ? k : cr + (k - cl);
        } else {
            const int i2 = idx - n1;
            const int rr = i2 / wc;
            c = cl + (i2 - rr * wc);
            r = rr < rt ? rr : rb + (rr - rt);
        }
        const int sx = (int)remap_index(P.style, (int64_t)xa + c, (int64_t)P.W);
        const int sy = (int)remap_index(P.style, (int64_t)ya + r, (int64_t)P.H);
        cell_src[idx] = (sx < 0 || sy < 0) ? -1 : sy * P.W + sx;
        cell_dst[idx] = (unsigned short)(r * S3_RWP + c);
    }
    if (tid == 0) {
        for (int i = 0; i < N; ++i) s3_mbar_init(full + 8 * i, 1);
        for (int i = 0; i < S3_NXF; ++i) s3_mbar_init(xfull + 8 * i, S3_NT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    const int qb = -2;
    auto locate_block = [&](int p0, int n) {
        if (tid < n) {
            const int p = p0 + tid;
            int which = -1, zc = 0;
            if (p >= 0 && p < in_planes) s3_locate(P, zin0 + p, which, zc);
            ptw[p & (S3_PT - 1)] = which;
            ptz[p & (S3_PT - 1)] = zc;
        }
    };
    locate_block(qb, S3_PTA);
    __syncthreads();

    bool lo_ready = P.flag_lo == nullptr, hi_ready = P.flag_hi == nullptr;     // thread 0 only
    auto wait_flag = [&](const unsigned char *f) {
        unsigned long long t0 = 0, t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*reinterpret_cast<const volatile unsigned char *>(f) != (unsigned char)P.epoch) {
            __nanosleep(64);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 5000000000ULL) asm volatile("trap;");          // 5 s: a copy that never arrives is an error, not a hang
        }
        __threadfence_system();
    };
    auto issue = [&](int p) {                   // one thread: TMA of input plane p into its ring buffer
        const int which = ptw[p & (S3_PT - 1)];
        if (which == 1 && !lo_ready) { wait_flag(P.flag_lo + (ya + in_rows > P.lo_early_rows ? 1 : 0)); lo_ready = true; }
        if (which == 2 && !hi_ready) { wait_flag(P.flag_hi); hi_ready = true; }
        const int zc = which < 0 ? P.own_n : ptz[p & (S3_PT - 1)];
        const void *map = which == 1 ? (const void *)&m_lo : which == 2 ? (const void *)&m_hi : (const void *)&m_own;
        const int b = p & (N - 1);
        s3_mbar_expect_tx(full + 8 * b, (unsigned)C::RAWBYTES);
        s3_tma_load3d(raw_sa + b * (RAWSZ * 4), map, full + 8 * b, xa, ya, zc);
    };
    auto fixup = [&](int p) {                   // all threads: the gather list of input plane p
        const int which = ptw[p & (S3_PT - 1)];
        const float *src = which < 0 ? nullptr
                                     : (which == 1 ? P.lo : which == 2 ? P.hi : P.own) + (long long)ptz[p & (S3_PT - 1)] * P.plane;
        float *dst = raw + (p & (N - 1)) * RAWSZ;
        for (int base = tid; base < ncell; base += 4 * S3_NT) {
            float v[4];
            int d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = base + k * S3_NT;
                d[k] = -1;
                if (idx < ncell) {
                    const int so = cell_src[idx];
                    d[k] = cell_dst[idx];
                    v[k] = (src != nullptr && so >= 0) ? __ldg(src + so) : P.fill;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (d[k] >= 0) dst[d[k]] = v[k];
        }
        if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };

    if (tma && tid == 0)
        for (int p = 0; p < min(S3_AHEAD, in_planes); ++p) issue(p);

    // this thread's column and 4 rows in stages y / z
    const int gx = x0 + lane, gy = y0 + 4 * warp;
    const int nrow = (gx >= 0 && gx < P.W) ? min(4, P.H - gy) : 0;                  // <= 0: nothing to store
    float *op = P.out + ((long long)(zo0 - P.own_first) + (qb - (Lz - 1))) * P.plane + (long long)gy * P.W + gx;
    const int yoff = (4 * warp) * S3_XFP + lane;

    float2 acc[2][LBZ];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int j = 0; j < LBZ; ++j) acc[o][j] = make_float2(0.f, 0.f);

    constexpr int NW = S3_NT / 32, XW = (C::RH + 7) / 8, XOFF = NW > XW ? NW - XW : 0;
    int wb = 0, wph = 0, ab = 0;                // xf slot / phase of plane q, xf slot of plane q + 1

    // ---- general interval (ramp-up, drain, volumes without TMA): every sub-step behind its run-time predicate -------------
    auto slow_interval = [&](const int q) {
        if (((q + 2) & (S3_PTB - 1)) == 0) locate_block(q + S3_PTA, S3_PTB);
        if (q >= 0) s3_mbar_wait(xfull + 8 * wb, wph);
        if (tma && tid == 0 && q + 2 + S3_AHEAD < in_planes) issue(q + 2 + S3_AHEAD);
        if (fix && q + 2 < in_planes) {
            const int p = q + 2;
            if (tma) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
            fixup(p);
        }
        if (q + 1 < in_planes && q + 1 >= 0) {
            const int p = q + 1;
            const int xw = (p & 1) ? warp - XOFF : warp;
            if (xw >= 0 && xw < XW) {
                if (tma && !fix) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
                s3_x_task<LXT, LYT, LZT>(P, raw_sa + (p & (N - 1)) * (RAWSZ * 4), xf_sa + ab * (XFSZ * 4), in_rows, Lx, xw, lane);
            }
            __syncwarp();
            if (lane == 0) s3_mbar_arrive(xfull + 8 * ab);
            ab = ab == S3_NXF - 1 ? 0 : ab + 1;
        }
        if (q < 0) {
            __syncthreads();                    // prologue: the gathers of planes 0 and 1 land before stage x reads them
        } else {
            float2 m[2];
            s3_y_task4<LXT, LYT, LZT>(P, xf_sa + (wb * XFSZ + yoff) * 4, m, Ly);
            s3_z_update4<LXT, LYT, LZT, CS>(P, acc, m, Lz, op, P.W, nrow, q >= Lz - 1 && nrow > 0);
            if (wb == S3_NXF - 1) { wb = 0; wph ^= 1; } else ++wb;
        }
        op += P.plane;
    };
    // ---- steady state: TMA issue, (patch,) stage x of plane q+1, stages y/z of plane q with a store: no predicates --------
    auto fast_interval = [&](auto fixc, const int q) {
        constexpr bool FIX = decltype(fixc)::value;
        if (((q + 2) & (S3_PTB - 1)) == 0) locate_block(q + S3_PTA, S3_PTB);
        s3_mbar_wait(xfull + 8 * wb, wph);
        if (tid == 0) issue(q + 2 + S3_AHEAD);
        if (FIX) {
            const int p = q + 2;
            s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
            fixup(p);
        }
        {
            const int p = q + 1;
            const int xw = (p & 1) ? warp - XOFF : warp;
            if (xw >= 0 && xw < XW) {
                if (!FIX) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
                s3_x_task<LXT, LYT, LZT>(P, raw_sa + (p & (N - 1)) * (RAWSZ * 4), xf_sa + ab * (XFSZ * 4), in_rows, Lx, xw, lane);
            }
            __syncwarp();
            if (lane == 0) s3_mbar_arrive(xfull + 8 * ab);
            ab = ab == S3_NXF - 1 ? 0 : ab + 1;
        }
        float2 m[2];
        s3_y_task4<LXT, LYT, LZT>(P, xf_sa + (wb * XFSZ + yoff) * 4, m, Ly);
        s3_z_update4<LXT, LYT, LZT, CS>(P, acc, m, Lz, op, P.W, nrow, nrow > 0);
        if (wb == S3_NXF - 1) { wb = 0; wph ^= 1; } else ++wb;
        op += P.plane;
    };

    // steady state = [max(Lz-1, 0), in_planes - 2 - AHEAD): stores on, planes q+1 and q+2+AHEAD exist
    const int q_fast0 = tma ? min(max(Lz - 1, 0), in_planes) : in_planes, q_fast1 = tma ? max(q_fast0, in_planes - 2 - S3_AHEAD) : in_planes;
    int q = -2;
    for (; q < q_fast0; ++q) slow_interval(q);
    if (fix) {
        for (; q < q_fast1; ++q) fast_interval(std::true_type{}, q);
    } else {
        for (; q < q_fast1; ++q) fast_interval(std::false_type{}, q);
    }
    for (; q < in_planes; ++q) slow_interval(q);
}

}  // namespace b2f
