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
    long long row_b, plane_b;      // the same pitches in bytes (W * 4, W * H * 4): the output pointer chains of v4
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
    // xy-filtered boundary planes (multi-GPU): planes whose stages x and y were run ahead of this launch (this rank's own
    // boundary planes, and the neighbours' through the exchange) — the march reads them in the z stage only.  xy_lo holds the
    // source planes [own_first - xlo_h, own_first + xlo_o), xy_hi the planes [own_first + own_n - xhi_o, own_first + own_n + xhi_h);
    // with them set, lo_n / hi_n only give the logical depth of the halos (lo / hi are not read).
    const float *xy_lo, *xy_hi;
    int xlo_h, xlo_o, xhi_o, xhi_h;
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

// ---- stage y of ONE plane: a thread owns one column x four consecutive rows (warp w = rows 4w..4w+3, lane = column) and
// reads 4 + Ly - 1 single floats of its column (LDS.32: 20 B of shared memory per voxel); the two row pairs are fed in the
// value-broadcast x tap-pair form of stage x ----------------------------------------------------------------------------------
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

}  // namespace b2f
