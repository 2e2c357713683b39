// stream3d_v3.cuh — round-2 form of the fused 3-D separable kernel (K1-3D): TWO planes per step, z stage paired ALONG z.
//
// Same tile (32 x 64 outputs), TMA ring, gather list and stage x as stream3d.cuh; what is new (and why — the numbers are in
// profiles/r2_stream3d_*):
//   * stages y / z: a thread owns ONE column x FOUR consecutive rows (warp w = rows 4w..4w+3, lane = column).  Stage y reads
//     single floats of its column (LDS.32: 20 B of shared memory per voxel instead of the 36 B of the 2 x 2 mapping).
//   * stage z pairs the two partial sums of ONE voxel that are adjacent in z: a step takes the xy-filtered values m0, m1 of
//     two consecutive planes and moves every pair one slot up, P[i] = P[i-1] + m0 * (k[2i-1], k[2i]) + m1 * (k[2i], k[2i+1]).
//     Every FFMA2 then has the form "one 32-bit value x a tap pair in a uniform register + a 64-bit partial-sum pair": a
//     single 64-bit register operand, so the value and the accumulator cannot collide in a register bank (the round-1 form —
//     value pair x scalar tap + accumulator pair — made ptxas re-copy the value for every other tap: ~28 MOV / IMAD.MOV per
//     warp and plane, the IMAD.MOVs on the FMA pipe).  The first and last tap of a voxel are scalar FFMAs, so the FMA-pipe
//     time stays Lz multiply-adds per voxel.
//   * one CTA-wide hand-off per TWO planes (split mbarrier: arrive after stage x of the next two planes, wait one step
//     later) over a 6-deep ring of x-filtered planes; loop counters, barrier parities, TMA issue and address updates are paid
//     once per two planes.
//   * stage x of the two planes of a step is 2 * XW = 20 warp-tasks for 16 warps; the warps that take a second task rotate
//     from step to step, one per scheduler (in round 1 warps 6-9 ran stage x on EVERY plane and all others waited for them).
//   * ramp-up / drain steps run a general step function; the steady-state step has no per-plane predicates, and border
//     patching is a compile-time flag of it.
#pragma once

#include "stream3d.cuh"

namespace b2f {

constexpr int S3V_NXF = 6;                        // ring of x-filtered planes (the planes of steps s-1, s, s+1)
constexpr int S3V_NPMAX = (S3_MAXTAPS - 1) / 2;   // partial-sum pairs per voxel

struct S3VTaps {                         // z taps as the pairs the paired form consumes (kk = taps right-aligned in K odd slots)
    float2 p0[S3V_NPMAX + 1];            // p0[i] = (kk[2i-1], kk[2i]),  kk[-1] = 0: multiplies m0
    float2 p1[S3V_NPMAX + 1];            // p1[i] = (kk[2i], kk[2i+1]),  kk[K]  = 0: multiplies m1
};

template <int LXT, int LYT, int LZT> struct S3VC : S3C<LXT, LYT, LZT> {
    typedef S3C<LXT, LYT, LZT> B;
    static constexpr int K = (B::LBZ & 1) ? B::LBZ : B::LBZ + 1;      // odd number of z slots
    static constexpr int NP = (K - 1) / 2;
    static constexpr size_t SMEM = sizeof(float) * (size_t)(S3_NRAW * B::RAWSZ + S3V_NXF * B::XFSZ) +
                                   (sizeof(int) + sizeof(short)) * B::NCELL + sizeof(int) * 2 * S3_PT + sizeof(uint64_t) * (S3_NRAW + 4);
};

// Stage y of BOTH planes of a step in one unrolled loop: four independent accumulator chains per thread.
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s3v_y_task4x2(const S3Params &P, const unsigned xa, const unsigned xb, float2 (&ma)[2], float2 (&mb)[2],
                                              const int Ly) {
    typedef S3C<LXT, LYT, LZT> C;
    ma[0] = ma[1] = mb[0] = mb[1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4 + C::LBY - 1; ++i) {
        if (LYT || i < 4 + Ly - 1) {
            const float sa = s3_lds32(xa + i * (S3_XFP * 4)), sb = s3_lds32(xb + i * (S3_XFP * 4));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = i - 2 * h;            // row pair h = outputs (2h, 2h+1): taps (j, j-1) of input row i
                if (j >= 0 && j <= C::LBY && (LYT || j <= Ly)) {
                    if (j == 0) { ma[h].x = fmaf(sa, P.ky[0], ma[h].x); mb[h].x = fmaf(sb, P.ky[0], mb[h].x); }
                    else if (j < C::LBY && (LYT || j < Ly)) { ma[h] = s3_fma2b(sa, P.kyp[j], ma[h]); mb[h] = s3_fma2b(sb, P.kyp[j], mb[h]); }
                    else if (LYT || j == Ly) { ma[h].y = fmaf(sa, P.ky[j - 1], ma[h].y); mb[h].y = fmaf(sb, P.ky[j - 1], mb[h].y); }
                }
            }
        }
    }
}

// non-blocking phase test (the result is consumed much later: the latency of the barrier unit is hidden)
__device__ __forceinline__ bool s3_mbar_test(unsigned b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    return ok != 0;
}

// One z step for the thread's 4 voxels.  P[v][i] = (A[2i+1], A[2i+2]) with A[t] the partial sum of the output that has
// received t of the K taps.  fin.y is completed by m0 (the EARLIER output plane), fin.x by m1.
template <int LXT, int LYT, int LZT, bool CS>
__device__ __forceinline__ void s3v_z_step(const S3VTaps &T, float2 (&P)[4][S3VC<LXT, LYT, LZT>::NP], const float (&m0)[4],
                                           const float (&m1)[4], const int Lz, float *__restrict__ op, const long long plane,
                                           const int W, const int nrow, const bool emit0, const bool emit1) {
    typedef S3VC<LXT, LYT, LZT> C;
    constexpr int NP = C::NP;
    float2 fin[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        fin[v] = s3_fma2b(m0[v], T.p0[NP], P[v][NP - 1]);
        fin[v].x = fmaf(m1[v], T.p1[NP].x, fin[v].x);               // (kk[K-1], 0): the later output only
    }
    // a pair whose taps are all leading zeros stays zero and is skipped: pair i is live iff 2i + 1 >= K - Lz
#pragma unroll
    for (int i = NP - 1; i >= 1; --i) {
        if (LZT || 2 * i + 1 >= C::K - Lz) {
#pragma unroll
            for (int v = 0; v < 4; ++v) P[v][i] = s3_fma2b(m1[v], T.p1[i], s3_fma2b(m0[v], T.p0[i], P[v][i - 1]));
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        float2 z;
        z.x = 0.f;
        z.y = m0[v] * T.p0[0].y;                                    // (0, kk[0]) * m0
        P[v][0] = s3_fma2b(m1[v], T.p1[0], z);
    }
    if (emit0) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (v < nrow) {
                if (CS) __stcs(op + (long long)v * W, fin[v].y); else op[(long long)v * W] = fin[v].y;
            }
    }
    if (emit1) {
        float *o1 = op + plane;
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (v < nrow) {
                if (CS) __stcs(o1 + (long long)v * W, fin[v].x); else o1[(long long)v * W] = fin[v].x;
            }
    }
}

template <int LXT, int LYT, int LZT, bool CS>
__global__ void __launch_bounds__(S3_NT, 1)
stream3d_kernel3(const __grid_constant__ S3Params P, const __grid_constant__ S3VTaps TZ, const __grid_constant__ CUtensorMap m_own,
                 const __grid_constant__ CUtensorMap m_lo, const __grid_constant__ CUtensorMap m_hi) {
    typedef S3VC<LXT, LYT, LZT> C;
    constexpr int TX = S3_TX, TY = S3_TY, RAWSZ = C::RAWSZ, XFSZ = C::XFSZ, N = S3_NRAW, NP = C::NP;

    extern __shared__ __align__(1024) float s3_smem[];
    float *raw = s3_smem;                       // N x RAWSZ
    float *xf = raw + N * RAWSZ;                // S3V_NXF x XFSZ
    int *cell_src = reinterpret_cast<int *>(xf + S3V_NXF * XFSZ);
    unsigned short *cell_dst = reinterpret_cast<unsigned short *>(cell_src + C::NCELL);
    int *ptw = reinterpret_cast<int *>(cell_dst + C::NCELL);
    int *ptz = ptw + S3_PT;
    // barriers: full + 8 b = TMA of raw buffer b landed; xfull + 8 i (i = step mod 3) = every warp is through stage x of the
    // two planes that step consumes
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
        for (int i = 0; i < 3; ++i) s3_mbar_init(xfull + 8 * i, S3_NT / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // plane sources: entry p & (S3_PT-1) describes input plane p; refilled S3_PTB planes at a time, S3_PTA planes ahead
    auto locate_block = [&](int p0, int n) {
        if (tid < n) {
            const int p = p0 + tid;
            int which = -1, zz = 0;
            if (p >= 0 && p < in_planes) s3_locate(P, zin0 + p, which, zz);
            ptw[p & (S3_PT - 1)] = which;
            ptz[p & (S3_PT - 1)] = zz;
        }
    };
    locate_block(-2, S3_PTA);                   // planes -2 .. PTA-3 (the first step refills the next block)
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
        const int zz = which < 0 ? P.own_n : ptz[p & (S3_PT - 1)];         // Fill(0) plane: out of range reads zero
        const void *map = which == 1 ? (const void *)&m_lo : which == 2 ? (const void *)&m_hi : (const void *)&m_own;
        const int b = p & (N - 1);
        s3_mbar_expect_tx(full + 8 * b, (unsigned)C::RAWBYTES);
        s3_tma_load3d(raw_sa + b * (RAWSZ * 4), map, full + 8 * b, xa, ya, zz);
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
    auto patch = [&](int p) {                   // all threads: wait for the TMA of plane p, then patch its border cells
        if (tma) s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
        fixup(p);
    };

    // prologue: planes 0..N-1 in flight.  Only thread 0 ever waits for a TMA to land: it checks the planes two steps before
    // stage x reads them and its next barrier arrival publishes that to the CTA (the other warps never touch the TMA
    // barriers: one try_wait latency less per warp-task).  Border tiles patch planes 0 and 1 before the first stage x.
    auto landed = [&](int p) { s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1); };
    if (tma && tid == 0) {
        for (int p = 0; p < min(N, in_planes); ++p) issue(p);
        for (int p = 0; p < min(4, in_planes); ++p) landed(p);
    }
    if (fix)
        for (int p = 0; p < min(2, in_planes); ++p) patch(p);
    __syncthreads();

    // this thread's column and 4 rows in stages y / z
    const int gx = x0 + lane, gy = y0 + 4 * warp;
    const int nrow = (gx >= 0 && gx < P.W) ? min(4, P.H - gy) : 0;                  // <= 0: nothing to store
    // step s (planes 2s, 2s+1) completes the output planes o = 2s - (Lz-1) [by m0] and o + 1 [by m1], relative to zo0
    float *op = P.out + ((long long)(zo0 - P.own_first) + (-2 - (Lz - 1))) * P.plane + (long long)gy * P.W + gx;   // step -1
    const int yoff = (4 * warp) * S3_XFP + lane;

    float2 acc[4][NP];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[v][i] = make_float2(0.f, 0.f);

    // Stage x of the two planes of a step = 2 * XW warp-tasks (20 for 17 taps) for the NXW = 15 warps 1..15 (warp 0 issues
    // and checks the TMAs instead); the warps that take a second task rotate by 2 * XW mod 15 from step to step.
    constexpr int NW = S3_NT / 32, XW = (C::RH + 7) / 8, NXW = NW - 1, XROT = (2 * XW) % NXW;
    static_assert(XW <= NXW && 2 * XW <= 2 * NXW, "stage x of two planes must fit two rounds of the x warps");
    int rot = 0;                                // rotation of this step's task round, 0 .. NXW-1
    int bi = 2, ph = 1;                         // barrier index of step s (s mod 3; step -1 counts as 2) and its phase
    bool ok = true;                             // early test of this step's barrier (made during the previous step)

    auto stage_x = [&](const int pa, const bool have_a, const bool have_b) {
        if (warp == 0) return;
        int k = warp - 1 - rot;
        if (k < 0) k += NXW;
        const int xn = bi == 2 ? 0 : 2 * bi + 2;                     // xf slots of planes pa, pa + 1
#pragma unroll
        for (int rnd = 0; rnd < 2; ++rnd, k += NXW) {
            if (k < XW) {
                if (have_a)
                    s3_x_task<LXT, LYT, LZT>(P, raw_sa + (pa & (N - 1)) * (RAWSZ * 4), xf_sa + xn * (XFSZ * 4), in_rows, Lx, k, lane);
            } else if (k < 2 * XW) {
                if (have_b)
                    s3_x_task<LXT, LYT, LZT>(P, raw_sa + ((pa + 1) & (N - 1)) * (RAWSZ * 4), xf_sa + (xn + 1) * (XFSZ * 4), in_rows, Lx,
                                             k - XW, lane);
            }
        }
    };
    auto arrive_next = [&]() {                  // this warp is through stage x of the next step's planes
        __syncwarp();
        if (lane == 0) s3_mbar_arrive(xfull + 8 * (bi == 2 ? 0 : bi + 1));
    };
    auto advance = [&]() {
        rot += XROT;
        if (rot >= NXW) rot -= NXW;
        if (bi == 2) { bi = 0; ph ^= 1; } else ++bi;
        op += 2 * P.plane;
    };
    // thread 0, after its arrival: the TMAs of planes p0+N, p0+N+1 go out (their ring buffers were read by stage x in the
    // previous step, which this step's barrier wait has seen complete)
    auto tma_work = [&](const int p0, const bool all) {
        if (tma && tid == 0) {
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int p = p0 + N + d;
                if (all || (p0 >= 0 && p < in_planes)) {
                    const int rel = zin0 + p - P.own_first;
                    if ((unsigned)rel < (unsigned)P.own_n) {                     // an owned plane: no table, no flag
                        const int b = p & (N - 1);
                        s3_mbar_expect_tx(full + 8 * b, (unsigned)C::RAWBYTES);
                        s3_tma_load3d(raw_sa + b * (RAWSZ * 4), &m_own, full + 8 * b, xa, ya, rel);
                    } else {
                        issue(p);
                    }
                }
            }
        }
    };
    // thread 0, BEFORE its arrival (warp 0 runs no stage x, so this sits in otherwise idle time): the planes the next step's
    // stage x reads have landed — they were issued two steps ago.  The arrival publishes it to the CTA.
    auto tma_check = [&](const int p0, const bool all) {
        if (tma && tid == 0) {
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const int p = p0 + 4 + d;
                if (all || p < in_planes) landed(p);
            }
        }
    };

    // ---- general step: every sub-step behind its run-time predicate (ramp-up, drain, volumes without TMA) --------------------
    auto slow_step = [&](const int s) {
        const int p0 = 2 * s;
        if (((p0 + 2) & (S3_PTB - 1)) == 0) locate_block(p0 + S3_PTA, S3_PTB);
        if (s >= 0 && !ok) s3_mbar_wait(xfull + 8 * bi, ph);
        ok = false;
        if (fix) {                                                  // the planes of the NEXT step's stage x
            if (p0 + 4 < in_planes) patch(p0 + 4);
            if (p0 + 5 < in_planes) patch(p0 + 5);
        }
        stage_x(p0 + 2, p0 + 2 < in_planes, p0 + 3 < in_planes);
        tma_check(p0, false);
        arrive_next();
        tma_work(p0, false);
        if (s >= 0) {
            const int o = p0 - (Lz - 1);
            float2 ma[2], mb[2];
            const unsigned xa0 = xf_sa + (2 * bi * XFSZ + yoff) * 4;
            if (p0 + 1 < in_planes) {
                s3v_y_task4x2<LXT, LYT, LZT>(P, xa0, xa0 + XFSZ * 4, ma, mb, Ly);
            } else {
                s3_y_task4<LXT, LYT, LZT>(P, xa0, ma, Ly);
                mb[0] = mb[1] = make_float2(0.f, 0.f);
            }
            const float m0[4] = {ma[0].x, ma[0].y, ma[1].x, ma[1].y}, m1[4] = {mb[0].x, mb[0].y, mb[1].x, mb[1].y};
            s3v_z_step<LXT, LYT, LZT, CS>(TZ, acc, m0, m1, Lz, op, P.plane, P.W, nrow, o >= 0 && o < nout && nrow > 0,
                                          o + 1 >= 0 && o + 1 < nout && nrow > 0);
        }
        advance();
    };
    // ---- steady state: both planes exist everywhere, both outputs are stored: no predicates ---------------------------------
    auto fast_step = [&](auto fixc, const int s) {
        constexpr bool FIX = decltype(fixc)::value;
        const int p0 = 2 * s;
        if (!ok) s3_mbar_wait(xfull + 8 * bi, ph);
        if (FIX) { patch(p0 + 4); patch(p0 + 5); }
        stage_x(p0 + 2, true, true);
        tma_check(p0, true);
        arrive_next();
        tma_work(p0, true);
        float2 ma[2], mb[2];
        const unsigned xa0 = xf_sa + (2 * bi * XFSZ + yoff) * 4;
        s3v_y_task4x2<LXT, LYT, LZT>(P, xa0, xa0 + XFSZ * 4, ma, mb, Ly);
        // the next step's barrier is tested here, a z stage ahead of its use: no warp sits out the barrier unit's latency
        ok = s3_mbar_test(xfull + 8 * (bi == 2 ? 0 : bi + 1), bi == 2 ? (ph ^ 1) : ph);
        const float m0[4] = {ma[0].x, ma[0].y, ma[1].x, ma[1].y}, m1[4] = {mb[0].x, mb[0].y, mb[1].x, mb[1].y};
        s3v_z_step<LXT, LYT, LZT, CS>(TZ, acc, m0, m1, Lz, op, P.plane, P.W, nrow, nrow > 0, nrow > 0);
        advance();
    };

    // Steps s = -1 .. nsteps-1.  Steady state: s >= 0; both outputs exist, 0 <= 2s-(Lz-1) and 2s+1-(Lz-1) <= nout-1; the planes
    // up to 2s+N+1 exist.
    const int nsteps = (in_planes + 1) >> 1;
    int s_fast0 = nsteps, s_fast1 = nsteps;
    if (tma && nout + Lz - 3 >= 0 && in_planes - N - 2 >= 0) {
        s_fast0 = min(Lz >> 1, nsteps);                                      // ceil((Lz-1)/2)
        s_fast1 = max(s_fast0, min(min((nout + Lz - 3) / 2 + 1, (in_planes - N - 2) / 2 + 1), nsteps));
    }
    int s = -1;
    for (; s < s_fast0; ++s) slow_step(s);
    while (s < s_fast1) {                       // the plane-source table is refilled every 8 steps, outside the inner loop
        if (((2 * s + 2) & (S3_PTB - 1)) == 0) locate_block(2 * s + S3_PTA, S3_PTB);
        const int e = min(s_fast1, (s + 1) | 7);
        if (fix) {
            for (; s < e; ++s) fast_step(std::true_type{}, s);
        } else {
            for (; s < e; ++s) fast_step(std::false_type{}, s);
        }
    }
    for (; s < nsteps; ++s) slow_step(s);
}

}  // namespace b2f
