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
constexpr int S3_NT = 32 * S3_NP * S3_GW;                  // 896 threads
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
    float kx[S3_MAXTAPS], ky[S3_MAXTAPS], kz[S3_MAXTAPS];
};

__device__ __forceinline__ float2 s3_fma2(float2 a, float k, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rc = *reinterpret_cast<unsigned long long *>(&c), rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(k));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
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

// LXT/LYT/LZT > 0: exact tap counts; 0: run-time count bounded by S3_MAXTAPS (uniform predicates)
template <int LXT, int LYT, int LZT>
__global__ void __launch_bounds__(S3_NT, 1) stream3d_kernel(const __grid_constant__ S3Params P) {
    constexpr int T = S3_T;
    constexpr int LBX = LXT ? LXT : S3_MAXTAPS, LBY = LYT ? LYT : S3_MAXTAPS, LBZ = LZT ? LZT : S3_MAXTAPS;
    constexpr int RH = T + LBY - 1;                        // raw / xf tile rows (compile-time bound)
    constexpr int RAWSZ = RH * S3_RWP, XFSZ = RH * S3_XFP, MIDSZ = T * T;
    constexpr int WINX = ((8 + LBX - 1 + 3) / 4) * 4;      // x window registers (whole 16-byte chunks), <= 24
    constexpr int WINZ = S3_RZ + LBZ - 1;                  // z window planes, <= 24
    static_assert(T + LBX - 1 <= S3_RWP && 8 * 3 + WINX <= S3_RWP, "raw pitch too small");
    static_assert(WINZ + 1 <= S3_RING, "ring too short");

    extern __shared__ __align__(16) float s3_smem[];
    float *raw = s3_smem;                       // S3_NRAW x RAWSZ
    float *xf = raw + S3_NRAW * RAWSZ;          // S3_NXF x XFSZ
    float *mid = xf + S3_NXF * XFSZ;            // S3_RING x MIDSZ
    int *yoff = reinterpret_cast<int *>(mid + S3_RING * MIDSZ);   // RH source rows (-1: Fill)

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

    for (int r = tid; r < in_rows; r += S3_NT) yoff[r] = (int)remap_index(P.style, (int64_t)y0 + P.kloy + r, (int64_t)P.H);
    __syncthreads();

    const int grp = warp / S3_GW, wr = warp % S3_GW;      // plane group, role index inside the group
    const bool is_x = wr < S3_XW, is_y = wr >= S3_XW && wr < S3_XW + S3_YW, is_z = wr >= S3_XW + S3_YW;
    const int yw = wr - S3_XW, zw = wr - S3_XW - S3_YW;

    // ---- loader state (y warps): 16-byte path when the whole window lies inside the row and is aligned --------------
    const int xa = x0 + P.klox;
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

    // ---- role geometry --------------------------------------------------------------------------------------------
    // x: a quarter-warp covers two rows x 32 columns (4 groups of 8), a warp 8 rows
    const int l8 = lane & 7, qw = lane >> 3;
    const int xg = l8 & 3, xr = 2 * qw + (l8 >> 2);
    // y: a half-warp covers 32 columns (16 pairs) of one 4-row group; 8 row groups over 4 warps x 2 half-warps
    const int yc = lane & 15, yg = 2 * yw + (lane >> 4);
    // z: a warp covers one row pair (2 rows x 16 column pairs); rows 8*st .. 8*st+7 form stagger class st
    const int zc = lane & 15, zrow_in_pair = lane >> 4;
    const int gx = x0 + 2 * zc;
    const int smode = gx >= P.W ? 0 : (gx + 1 >= P.W ? 1 : (P.vec_out ? 3 : 2));

    // Phase q: group g runs stage x on plane 2q+g, stage y on plane 2q-2+g, stage z on the block whose last input
    // plane is 2q-4+g, and prefetches plane 2q+4+g.
    if (is_y) { load_plane(grp, grp % S3_NRAW); load_plane(S3_NP + grp, (S3_NP + grp) % S3_NRAW); }

    const int nphase = (in_planes + S3_RZ + 3) / S3_NP + 2;      // the last phases only drain partial z blocks
    for (int q = 0; q < nphase; ++q) {
        if (is_y) s3_wait<1>();
        __syncthreads();
        if (is_x) {
            // ---- stage x of plane px: raw[px % NRAW] -> xf[px % NXF] ------------------------------------------------
            const int px = S3_NP * q + grp;
            if (px < in_planes) {
                const float *rb = raw + (px % S3_NRAW) * RAWSZ;
                float *xb = xf + (px % S3_NXF) * XFSZ;
                for (int row = 8 * wr + xr; row < in_rows; row += 8 * S3_XW) {
                    const float *src = rb + row * S3_RWP + 8 * xg;
                    float v[WINX];
#pragma unroll
                    for (int i = 0; i < WINX; i += 4) {
                        const float4 t = *reinterpret_cast<const float4 *>(src + i);
                        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
                    }
                    float2 a[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) a[k] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < LBX; ++j) {
                        if (LXT || j < Lx) {
                            const float k = P.kx[j];
                            // even taps read aligned register pairs (packed FFMA2); odd taps would need two moves per
                            // pair, so they issue as two scalar FFMAs: same FP32-pipe cost, no extra instructions
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                if (j % 2 == 0) {
                                    a[c] = s3_fma2(make_float2(v[2 * c + j], v[2 * c + j + 1]), k, a[c]);
                                } else {
                                    a[c].x = fmaf(v[2 * c + j], k, a[c].x);
                                    a[c].y = fmaf(v[2 * c + j + 1], k, a[c].y);
                                }
                            }
                        }
                    }
                    float *d = xb + row * S3_XFP + 8 * xg;
                    *reinterpret_cast<float4 *>(d) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
                    *reinterpret_cast<float4 *>(d + 4) = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
                }
            }
        } else if (is_y) {
            const int pl = S3_NP * (q + 2) + grp;
            load_plane(pl, pl % S3_NRAW);
            // ---- stage y of plane py: xf[py % NXF] -> mid[py % RING] --------------------------------------------------
            const int py = S3_NP * (q - 1) + grp;
            if (py >= 0 && py < in_planes) {
                const float *xb = xf + (py % S3_NXF) * XFSZ + (4 * yg) * S3_XFP + 2 * yc;
                float2 m[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) m[o] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4 + LBY - 1; ++i) {
                    if (LYT || i < 4 + Ly - 1) {
                        const float2 s = *reinterpret_cast<const float2 *>(xb + i * S3_XFP);
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            const int j = i - o;
                            if (j >= 0 && j < LBY && (LYT || j < Ly)) m[o] = s3_fma2(s, P.ky[j], m[o]);
                        }
                    }
                }
                float *mb = mid + (py & (S3_RING - 1)) * MIDSZ + (4 * yg) * T + 2 * yc;
#pragma unroll
                for (int o = 0; o < 4; ++o) *reinterpret_cast<float2 *>(mb + o * T) = m[o];
            }
        } else if (is_z) {
            // ---- stage z: the block of RZ output planes starting at o0 (chunk-local) became complete with plane 2q-4+g ----
            const int o0 = S3_NP * (q - 2) + grp - (S3_RZ - 1) - (Lz - 1);   // its last input plane is o0 + RZ-1 + Lz-1
            const int st = ((o0 % S3_RZ) + S3_RZ) % S3_RZ;  // stagger class whose blocks start at o0
            const int olo = max(o0, 0), ohi = min(o0 + S3_RZ, nout);
            if (olo < ohi) {
                const int row = 8 * st + 2 * zw + zrow_in_pair;            // row of the tile
                // ring walk in byte offsets: one add and one mask per plane (the ring is a power of two long)
                constexpr unsigned MIDB = MIDSZ * 4u, RINGB = S3_RING * MIDB;
                const unsigned off0 = (unsigned)(o0 & (S3_RING - 1)) * MIDB + (unsigned)(row * T + 2 * zc) * 4u;
                const char *mbase = reinterpret_cast<const char *>(mid);
                float2 w[WINZ];
#pragma unroll
                for (int i = 0; i < WINZ; ++i)
                    if (LZT || i < S3_RZ + Lz - 1)
                        w[i] = *reinterpret_cast<const float2 *>(mbase + ((off0 + i * MIDB) & (RINGB - 1)));
                const int gy = y0 + row;
                if (gy < P.H && smode != 0) {
                    float *op = P.out + (long long)(zo0 - P.own_first + o0) * P.plane + (long long)gy * P.W + gx;
                    float2 acc[S3_RZ];
#pragma unroll
                    for (int o = 0; o < S3_RZ; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < LBZ; ++j) {
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
            }
        }
    }
    if (is_y) s3_wait<0>();
}

}  // namespace b2f
