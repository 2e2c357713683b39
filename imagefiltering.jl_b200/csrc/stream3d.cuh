// stream3d.cuh — K1-3D: fused 3-D separable FIR (x, y, z stages in ONE launch), Float32, slab-aware.
//
// Replaces, for a cascade of three 1-D factors on axes 1,2,3 (KernelFactors.gaussian((s,s,s)) & friends), the
// reference's padarray + three full-volume passes through padded temporaries (src/imfilter.jl:321-341, 385-395,
// 438-446, loops :724-739): every input voxel is read from HBM once, every output voxel written once.
//
//   * a CTA owns a 32 x 32 tile of the xy-plane and MARCHES along z over its chunk of planes, one plane per barrier
//     interval ("phase"); its 12 warps are specialised, all three roles run concurrently on different planes:
//       - loader (the y warps): plane q+2 -> cp.async (LDGSTS, 16-byte chunks on x-interior tiles) into a 3-deep ring
//         of raw tiles (tile + halo).  This is where the border remap / Fill of src/border.jl:564-590 is applied, in
//         x, y AND z, i.e. "pad the input once", exactly the reference's semantics;
//       - x warps (6): stage x of plane q: 8 adjacent outputs per thread from a 24-value register window (6 LDS.128;
//         the tile pitches are an odd number of 16-byte chunks, so a quarter-warp spanning two rows hits every bank
//         once) -> xf tile (double buffered);
//       - y warps (4): stage y of plane q-1: a thread owns 2 columns x 4 rows, 20 LDS.64 feed 4 register accumulators
//         per column pair -> slot (q-1) % 32 of a ring of xy-filtered planes in shared memory;
//       - z warps (2): stage z, register-blocked along z: a thread produces 8 consecutive output planes of its column
//         pair from a 24-plane register window read out of the ring (24 LDS.64 for 8 x 2 outputs) and stores them
//         (one 8-byte store per thread and plane, 128-byte rows).  The 16 row pairs of the tile take turns, two per
//         phase (their 8-plane blocks are staggered), so every phase carries the same work;
//   * every tap loop is unrolled with static register indices and ascending tap order (the reference's order); all
//     multiply-adds are packed FFMA2 (fma.rn.f32x2: two voxels per instruction, the tap broadcast from a uniform
//     register): the kernel sits at the FP32 ridge (51 FMA per 8 B of HBM traffic), and halving the FMA issue slots
//     is what leaves room for the LDS / LDGSTS / STG stream;
//   * one __syncthreads per plane; about 150 KB of shared memory, 1 CTA (12 warps) per SM.
//
// Slab form (multi-GPU): the planes of the last axis may live in three buffers — this rank's owned planes
// plus `lo`/`hi` halo planes, which are either receive buffers or PEER memory of the neighbouring GPU mapped over
// NVLink (the kernel then performs the halo exchange itself, by P2P loads, fused with the filter).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace b2f {

constexpr int S3_T = 32;           // tile edge (outputs): 32 x 32
constexpr int S3_MAXTAPS = 17;
constexpr int S3_RZ = 4;           // output planes per z block
constexpr int S3_RING = 32;        // xy-filtered planes kept in shared memory (>= RZ + MAXTAPS - 1 + 1; power of two)
constexpr int S3_RWP = 52;         // raw tile pitch in floats: 13 chunks of 16 B (odd)
constexpr int S3_XFP = 36;         // x-filtered tile pitch: 9 chunks (odd)
constexpr int S3_XW = 6, S3_YW = 4, S3_ZW = 4;             // warps per role
constexpr int S3_NP = 2;                                   // planes per phase: two warp groups, one plane each
constexpr int S3_GW = S3_XW + S3_YW + S3_ZW;               // warps per group (14)
constexpr int S3_NT = 32 * (S3_NP * S3_GW + 1);            // 928 threads: two worker groups + the TMA producer warp
constexpr int S3_NRAW = 3 * S3_NP, S3_NXF = 2 * S3_NP;     // raw / xf tile buffers

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
    int vec_in, vec_out;           // 16-byte loads / 8-byte stores are aligned
    int use_tma;                   // the tensor maps are valid: interior tiles take the pipelined path
    float kx[S3_MAXTAPS], ky[S3_MAXTAPS], kz[S3_MAXTAPS];
    float2 kxp[S3_MAXTAPS];        // kxp[j] = (kx[j], kx[j-1]): the taps one input value carries to two adjacent outputs
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
__device__ __forceinline__ void s3_cp16(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void s3_cp4(float *dst, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void s3_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void s3_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// source of global plane index zi: halo buffers are matched on the LOGICAL index first (a circular wrap arrives as
// an ordinary halo), then the border remap is applied in global coordinates.  nullptr = Fill plane.
__device__ __forceinline__ const float *s3_plane(const S3Params &P, int zi) {
    const int rel = zi - P.own_first;
    if ((unsigned)rel < (unsigned)P.own_n) return P.own + (long long)rel * P.plane;
    if (rel < 0 && rel >= -P.lo_n) return P.lo + (long long)(rel + P.lo_n) * P.plane;
    if (rel >= P.own_n && rel < P.own_n + P.hi_n) return P.hi + (long long)(rel - P.own_n) * P.plane;
    const int g = (int)remap_index(P.style, (int64_t)zi, (int64_t)P.Zg);
    if (g < 0) return nullptr;
    const int r2 = g - P.own_first;
    if ((unsigned)r2 < (unsigned)P.own_n) return P.own + (long long)r2 * P.plane;
    if (r2 < 0) return P.lo + (long long)(r2 + P.lo_n) * P.plane;      // host validated that it is present
    return P.hi + (long long)(r2 - P.own_n) * P.plane;
}

// ---- mbarrier / TMA primitives (sm_90+ PTX) ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned s3_sa(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s3_mbar_init(uint64_t *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_sa(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void s3_mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s3_sa(b)) : "memory");
}
__device__ __forceinline__ void s3_mbar_expect_tx(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_sa(b)), "r"(bytes) : "memory");
}
// Blocks in hardware until the phase with `parity` completes (the suspend-time hint keeps a waiting warp off the issue
// port instead of spinning: re-polling warps were taking a quarter of all issue slots).
__device__ __forceinline__ void s3_mbar_wait(uint64_t *b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(s3_sa(b)), "r"(parity), "r"(1000000u) : "memory");
        if (!ok) __nanosleep(100);
    } while (!ok);
}
__device__ __forceinline__ void s3_tma_load3d(float *dst, const void *tmap, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(s3_sa(dst)), "l"(tmap), "r"(s3_sa(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void s3_tma_prefetch3d(const void *tmap, int c0, int c1, int c2) {   // global -> L2 only
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// compile-time geometry shared by both execution paths
template <int LXT, int LYT, int LZT> struct S3C {
    static constexpr int T = S3_T;
    static constexpr int LBX = LXT ? LXT : S3_MAXTAPS, LBY = LYT ? LYT : S3_MAXTAPS, LBZ = LZT ? LZT : S3_MAXTAPS;
    static constexpr int RH = T + LBY - 1;                        // raw / xf tile rows (compile-time bound)
    static constexpr int RAWSZ = RH * S3_RWP, XFSZ = RH * S3_XFP, MIDSZ = T * T;
    static constexpr int WINX = ((8 + LBX - 1 + 3) / 4) * 4;      // x window registers (whole 16-byte chunks), <= 24
    static constexpr int WINZ = S3_RZ + LBZ - 1;                  // z window planes, <= 20
    static constexpr int NBAR = 2 * S3_NRAW + 2 * S3_NXF + 2 * S3_RING;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(S3_NRAW * RAWSZ + S3_NXF * XFSZ + S3_RING * MIDSZ) +
                                   sizeof(int) * (size_t)((RH + 1) & ~1) + sizeof(uint64_t) * NBAR;
    static_assert(T + LBX - 1 <= S3_RWP && 8 * 3 + WINX <= S3_RWP, "raw pitch too small");
    static_assert(WINZ + 2 * S3_NP <= S3_RING, "ring too short");
};

// ---- stage x: 8 rows of the tile per warp.  A quarter-warp covers two rows x 32 columns (4 groups of 8 outputs) ----
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_x_task(const S3Params &P, const float *__restrict__ rb, float *__restrict__ xb,
                                          const int row0, const int in_rows, const int Lx, const int lane) {
    typedef S3C<LXT, LYT, LZT> C;
    const int l8 = lane & 7, qw = lane >> 3;
    const int xg = l8 & 3, row = row0 + 2 * qw + (l8 >> 2);
    if (row >= in_rows) return;
    const float *src = rb + row * S3_RWP + 8 * xg;
    float v[C::WINX];
#pragma unroll
    for (int i = 0; i < C::WINX; i += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(src + i);
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
    float *d = xb + row * S3_XFP + 8 * xg;
    *reinterpret_cast<float4 *>(d) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
    *reinterpret_cast<float4 *>(d + 4) = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
}

// ---- stage y: a half-warp covers 32 columns (16 pairs) of one 4-row group; 8 row groups over 4 warps ----------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_y_task(const S3Params &P, const float *__restrict__ xfb, float *__restrict__ slot,
                                          const int yw, const int Ly, const int lane) {
    typedef S3C<LXT, LYT, LZT> C;
    const int yc = lane & 15, yg = 2 * yw + (lane >> 4);
    const float *xb = xfb + (4 * yg) * S3_XFP + 2 * yc;
    float2 m[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) m[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4 + C::LBY - 1; ++i) {
        if (LYT || i < 4 + Ly - 1) {
            const float2 s = *reinterpret_cast<const float2 *>(xb + i * S3_XFP);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int j = i - o;
                if (j >= 0 && j < C::LBY && (LYT || j < Ly)) m[o] = s3_fma2(s, P.ky[j], m[o]);
            }
        }
    }
    float *mb = slot + (4 * yg) * C::T + 2 * yc;
#pragma unroll
    for (int o = 0; o < 4; ++o) *reinterpret_cast<float2 *>(mb + o * C::T) = m[o];
}

// ---- stage z: the block of RZ output planes starting at (chunk-local) o0; a warp covers one row pair (2 rows x 16 column
// pairs); rows 8*st .. 8*st+7 form the stagger class st = o0 mod RZ, 4 warps cover it --------------------------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3_z_task(const S3Params &P, const float *__restrict__ mid, const int o0, const int zw,
                                          const int x0, const int y0, const int zo0, const int nout, const int Lz,
                                          const int lane) {
    typedef S3C<LXT, LYT, LZT> C;
    const int olo = max(o0, 0), ohi = min(o0 + S3_RZ, nout);
    if (olo >= ohi) return;
    const int st = o0 & (S3_RZ - 1);
    const int zc = lane & 15;
    const int row = 8 * st + 2 * zw + (lane >> 4);
    const int gx = x0 + 2 * zc, gy = y0 + row;
    const int smode = gx >= P.W ? 0 : (gx + 1 >= P.W ? 1 : (P.vec_out ? 3 : 2));   // 3: 8-byte store, 2: two scalars, 1: one
    // ring walk in byte offsets: one add and one mask per plane (the ring is a power of two long)
    constexpr unsigned MIDB = C::MIDSZ * 4u, RINGB = S3_RING * MIDB;
    const unsigned off0 = (unsigned)(o0 & (S3_RING - 1)) * MIDB + (unsigned)(row * C::T + 2 * zc) * 4u;
    const char *mbase = reinterpret_cast<const char *>(mid);
    float2 w[C::WINZ];
#pragma unroll
    for (int i = 0; i < C::WINZ; ++i)
        if (LZT || i < S3_RZ + Lz - 1) w[i] = *reinterpret_cast<const float2 *>(mbase + ((off0 + i * MIDB) & (RINGB - 1)));
    if (gy >= P.H || smode == 0) return;
    float *op = P.out + (long long)(zo0 - P.own_first + o0) * P.plane + (long long)gy * P.W + gx;
    float2 acc[S3_RZ];
#pragma unroll
    for (int o = 0; o < S3_RZ; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < C::LBZ; ++j) {
        if (LZT || j < Lz) {
            const float k = P.kz[j];
#pragma unroll
            for (int o = 0; o < S3_RZ; ++o) acc[o] = s3_fma2(w[o + j], k, acc[o]);
        }
    }
    if (smode == 3 && olo == o0 && ohi == o0 + S3_RZ) {       // whole block, aligned rows: the common case
#pragma unroll
        for (int o = 0; o < S3_RZ; ++o) *reinterpret_cast<float2 *>(op + (long long)o * P.plane) = acc[o];
    } else {
#pragma unroll
        for (int o = 0; o < S3_RZ; ++o) {
            if (o0 + o >= olo && o0 + o < ohi) {
                float *qp = op + (long long)o * P.plane;
                if (smode == 3) {
                    *reinterpret_cast<float2 *>(qp) = acc[o];
                } else {
                    qp[0] = acc[o].x;
                    if (smode == 2) qp[1] = acc[o].y;
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
    constexpr int T = C::T, RAWSZ = C::RAWSZ, XFSZ = C::XFSZ, MIDSZ = C::MIDSZ, RH = C::RH;

    extern __shared__ __align__(1024) float s3_smem[];
    float *raw = s3_smem;                       // S3_NRAW x RAWSZ
    float *xf = raw + S3_NRAW * RAWSZ;          // S3_NXF x XFSZ
    float *mid = xf + S3_NXF * XFSZ;            // S3_RING x MIDSZ
    int *yoff = reinterpret_cast<int *>(mid + S3_RING * MIDSZ);   // RH source rows (-1: Fill)
    uint64_t *bars = reinterpret_cast<uint64_t *>(yoff + ((RH + 1) & ~1));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Lx = LXT ? LXT : P.Lx, Ly = LYT ? LYT : P.Ly, Lz = LZT ? LZT : P.Lz;
    const int bid = blockIdx.x;
    const int tx = bid % P.ntx, ty = (bid / P.ntx) % P.nty, ch = bid / (P.ntx * P.nty);
    const int x0 = tx * T, y0 = ty * T;
    const int in_cols = T + Lx - 1, in_rows = T + Ly - 1;
    const int zo0 = P.own_first + ch * P.zchunk;                         // first output plane of this chunk (global)
    const int nout = min(P.zchunk, P.own_first + P.own_n - zo0);
    const int in_planes = nout + Lz - 1;
    const int zin0 = zo0 + P.kloz;                                       // global index of input plane p = 0
    const int xa = x0 + P.klox, ya = y0 + P.kloy;

    const int grp = warp / S3_GW, wr = warp % S3_GW;      // plane group (2 = the producer warp), role index inside the group
    const bool worker = grp < S3_NP;
    const bool is_x = worker && wr < S3_XW, is_y = worker && wr >= S3_XW && wr < S3_XW + S3_YW,
               is_z = worker && wr >= S3_XW + S3_YW;
    const int yw = wr - S3_XW, zw = wr - S3_XW - S3_YW;

    // ================================================================================================================
    // Pipelined path (tiles whose input window needs no border remap in x / y): TMA loads, mbarrier hand-offs between
    // the roles, no CTA-wide barrier in the plane loop.
    //   producer lane : plane p -> raw[p % NRAW]         (cp.async.bulk.tensor; out-of-range rows/planes read as zero)
    //   x warps (grp g = p % 2): raw -> xf[p % NXF];  y warps: xf -> ring slot p % RING;  z warps: ring -> out
    // ================================================================================================================
    if (P.use_tma && xa >= 0 && xa + in_cols <= P.W && ya >= 0 && ya + in_rows <= P.H) {
        uint64_t *raw_full = bars, *raw_empty = raw_full + S3_NRAW, *xf_full = raw_empty + S3_NRAW,
                 *xf_empty = xf_full + S3_NXF, *ring_full = xf_empty + S3_NXF, *zdone = ring_full + S3_RING;
        if (tid == 0) {
            for (int i = 0; i < S3_NRAW; ++i) { s3_mbar_init(raw_full + i, 1); s3_mbar_init(raw_empty + i, S3_XW); }
            for (int i = 0; i < S3_NXF; ++i) { s3_mbar_init(xf_full + i, S3_XW); s3_mbar_init(xf_empty + i, S3_YW); }
            for (int i = 0; i < S3_RING; ++i) { s3_mbar_init(ring_full + i, S3_YW); s3_mbar_init(zdone + i, S3_ZW); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int s_first = -(S3_RZ - 1);                   // first block start that owns a valid output
        const int nblocks = nout - s_first;                 // block starts s_first .. nout-1
        if (!worker) {
            if (lane == 0) {
                // source of plane p: own / halo buffers matched on the logical index, else the border remap
                auto locate = [&](int p, const void *&map, int &zc) {
                    const int zi = zin0 + p;
                    int rel = zi - P.own_first;
                    map = &m_own;
                    if ((unsigned)rel < (unsigned)P.own_n) { zc = rel; return; }
                    if (rel < 0 && rel >= -P.lo_n) { map = &m_lo; zc = rel + P.lo_n; return; }
                    if (rel >= P.own_n && rel < P.own_n + P.hi_n) { map = &m_hi; zc = rel - P.own_n; return; }
                    const int g = (int)remap_index(P.style, (int64_t)zi, (int64_t)P.Zg);
                    rel = g - P.own_first;
                    if (g < 0) { zc = P.own_n; }                                   // Fill(0): out of range reads zero
                    else if ((unsigned)rel < (unsigned)P.own_n) { zc = rel; }
                    else if (rel < 0) { map = &m_lo; zc = rel + P.lo_n; }
                    else { map = &m_hi; zc = rel - P.own_n; }
                };
                // the shared-memory ring holds S3_NRAW planes (about one DRAM latency of work); planes further ahead are
                // pulled into L2 by TMA prefetches, so the loads that fill the ring are L2 hits
                constexpr int PF = 10;
                const void *map;
                int zc;
                for (int p = 0; p < min(PF, in_planes); ++p) { locate(p, map, zc); s3_tma_prefetch3d(map, xa, ya, zc); }
                for (int p = 0; p < in_planes; ++p) {
                    const int b = p % S3_NRAW, k = p / S3_NRAW;
                    if (p + PF < in_planes) { locate(p + PF, map, zc); s3_tma_prefetch3d(map, xa, ya, zc); }
                    s3_mbar_wait(raw_empty + b, (k & 1) ^ 1);
                    locate(p, map, zc);
                    s3_mbar_expect_tx(raw_full + b, (unsigned)(RAWSZ * sizeof(float)));
                    s3_tma_load3d(raw + b * RAWSZ, map, raw_full + b, xa, ya, zc);
                }
            }
        } else if (is_x) {
            for (int p = grp; p < in_planes; p += S3_NP) {
                const int b = p % S3_NRAW, xbuf = p % S3_NXF;
                s3_mbar_wait(raw_full + b, (p / S3_NRAW) & 1);
                s3_mbar_wait(xf_empty + xbuf, ((p / S3_NXF) & 1) ^ 1);
                s3_x_task<LXT, LYT, LZT>(P, raw + b * RAWSZ, xf + xbuf * XFSZ, 8 * wr, in_rows, Lx, lane);
                __syncwarp();
                if (lane == 0) { s3_mbar_arrive(xf_full + xbuf); s3_mbar_arrive(raw_empty + b); }
            }
        } else if (is_y) {
            for (int p = grp; p < in_planes; p += S3_NP) {
                const int xbuf = p % S3_NXF, slot = p & (S3_RING - 1);
                // the slot still holds plane p - RING: every z block that reads it (starts <= p - RING) must be done;
                // blocks of one group finish in order, so the newest block of each group is enough
                const int bw = p - S3_RING - s_first;
                if (bw >= 0) s3_mbar_wait(zdone + (bw & (S3_RING - 1)), (bw / S3_RING) & 1);
                if (bw >= 1) s3_mbar_wait(zdone + ((bw - 1) & (S3_RING - 1)), ((bw - 1) / S3_RING) & 1);
                s3_mbar_wait(xf_full + xbuf, (p / S3_NXF) & 1);
                s3_y_task<LXT, LYT, LZT>(P, xf + xbuf * XFSZ, mid + slot * MIDSZ, yw, Ly, lane);
                __syncwarp();
                if (lane == 0) { s3_mbar_arrive(ring_full + slot); s3_mbar_arrive(xf_empty + xbuf); }
            }
        } else if (is_z) {
            for (int b = grp; b < nblocks; b += S3_NP) {
                const int o0 = s_first + b;
                const int pn = min(o0 + S3_RZ + Lz - 2, in_planes - 1);      // newest plane the block reads
                s3_mbar_wait(ring_full + (pn & (S3_RING - 1)), (pn / S3_RING) & 1);
                if (pn >= 1) s3_mbar_wait(ring_full + ((pn - 1) & (S3_RING - 1)), ((pn - 1) / S3_RING) & 1);
                s3_z_task<LXT, LYT, LZT>(P, mid, o0, zw, x0, y0, zo0, nout, Lz, lane);
                __syncwarp();
                if (lane == 0) s3_mbar_arrive(zdone + (b & (S3_RING - 1)));
            }
        }
        return;
    }

    // ================================================================================================================
    // Barrier path (border tiles, Fill(v != 0), unaligned arrays): cp.async loads through the border remap, one
    // __syncthreads per phase of two planes.
    // ================================================================================================================
    for (int r = tid; r < in_rows; r += S3_NT) yoff[r] = (int)remap_index(P.style, (int64_t)ya + r, (int64_t)P.H);
    __syncthreads();

    // ---- loader state (y warps): 16-byte path when the whole window lies inside the row and is aligned --------------
    const int in_cols4 = (in_cols + 3) & ~3;
    const bool vec = P.vec_in && xa >= 0 && xa + in_cols4 <= P.W && (xa & 3) == 0;
    const int nchunk = in_cols4 >> 2;                     // <= 12
    // vector path: lane -> (row parity, chunk): 2 rows x 16 chunk slots per warp iteration
    const int lrow = lane >> 4, lchk = lane & 15;
    int xo[2] = {-2, -2};                                 // scalar path: source columns of this lane (-1 Fill, -2 none)
    if (is_y && !vec) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c = lane + 32 * k;
            if (c < in_cols) xo[k] = (int)remap_index(P.style, (int64_t)xa + c, (int64_t)P.W);
        }
    }
    // vector path: this lane's (row, chunk) slots are the same for every plane: keep their source offsets in registers
    constexpr int NLD = (RH + 2 * S3_YW - 1) / (2 * S3_YW);       // rows per lane and plane (6 for 17 taps)
    int goff[NLD];                                                // element offset inside a plane; -1: Fill; -2: none
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int r = 2 * yw + lrow + 2 * S3_YW * i;
        goff[i] = -2;
        if (is_y && vec && r < in_rows && lchk < nchunk) {
            const int yo = yoff[r];
            goff[i] = yo < 0 ? -1 : yo * P.W + xa + 4 * lchk;
        }
    }
    auto load_plane = [&](int p, int buf) {
        if (p < in_planes) {
            const float *src = s3_plane(P, zin0 + p);
            float *dst = raw + buf * RAWSZ;
            if (vec) {
                float *d = dst + (2 * yw + lrow) * S3_RWP + 4 * lchk;
#pragma unroll
                for (int i = 0; i < NLD; ++i) {
                    if (goff[i] != -2) {
                        if (src != nullptr && goff[i] >= 0) s3_cp16(d, src + goff[i]);
                        else *reinterpret_cast<float4 *>(d) = make_float4(P.fill, P.fill, P.fill, P.fill);
                    }
                    d += 2 * S3_YW * S3_RWP;
                }
            } else {
                for (int r = yw; r < in_rows; r += S3_YW) {
                    const int yo = yoff[r];
                    const float *srow = src + (long long)(yo < 0 ? 0 : yo) * P.W;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        if (xo[k] != -2) {
                            float *d = dst + r * S3_RWP + lane + 32 * k;
                            if (src != nullptr && yo >= 0 && xo[k] >= 0) s3_cp4(d, srow + xo[k]);
                            else *d = P.fill;
                        }
                    }
                }
            }
        }
        s3_commit();
    };

    // Phase q: group g runs stage x on plane 2q+g, stage y on plane 2q-2+g, stage z on the block whose last input
    // plane is 2q-4+g, and prefetches plane 2q+4+g.
    if (is_y) { load_plane(grp, grp % S3_NRAW); load_plane(S3_NP + grp, (S3_NP + grp) % S3_NRAW); }

    const int nphase = (in_planes + S3_RZ + 3) / S3_NP + 2;      // the last phases only drain partial z blocks
    for (int q = 0; q < nphase; ++q) {
        if (is_y) s3_wait<1>();
        __syncthreads();
        if (is_x) {
            const int px = S3_NP * q + grp;
            if (px < in_planes)
                s3_x_task<LXT, LYT, LZT>(P, raw + (px % S3_NRAW) * RAWSZ, xf + (px % S3_NXF) * XFSZ, 8 * wr, in_rows, Lx, lane);
        } else if (is_y) {
            const int pl = S3_NP * (q + 2) + grp;
            load_plane(pl, pl % S3_NRAW);
            const int py = S3_NP * (q - 1) + grp;
            if (py >= 0 && py < in_planes)
                s3_y_task<LXT, LYT, LZT>(P, xf + (py % S3_NXF) * XFSZ, mid + (py & (S3_RING - 1)) * MIDSZ, yw, Ly, lane);
        } else if (is_z) {
            const int o0 = S3_NP * (q - 2) + grp - (S3_RZ - 1) - (Lz - 1);   // its last input plane is o0 + RZ-1 + Lz-1
            s3_z_task<LXT, LYT, LZT>(P, mid, o0, zw, x0, y0, zo0, nout, Lz, lane);
        }
    }
    if (is_y) s3_wait<0>();
}

}  // namespace b2f
