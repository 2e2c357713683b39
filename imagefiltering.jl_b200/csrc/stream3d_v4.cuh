// stream3d_v4.cuh — the fused 3-D separable kernel (K1-3D), round-2 final form.  Same algorithm, tile, TMA ring, hand-off
// protocol and z pairing as stream3d_v3.cuh; what changed is the instruction stream AROUND the multiply-adds (the v3 profile,
// profiles/r2_stream3d_c5_296planes_summary.txt: FFMA2 only 37 % of issued instructions, the FMA pipe 51 % busy):
//   * stage x is ONE inlined body behind a two-trip loop over the warp's task numbers (v3 inlined four copies — two rounds x
//     two planes — 17 KB of loop code); the lane-dependent parts of its shared-memory addresses are computed once per CTA;
//   * a finished pair of planes leaves through pointer chains with the row / plane pitches as kernel parameters in bytes
//     (v3 re-derived W * 4, its sign extension and 64-bit products in every step: ~35 instructions for 8 stores);
//   * interior lanes (all 4 rows inside the volume) store without predicates.
#pragma once

#include "stream3d_v3.cuh"

namespace b2f {

// ---- stage x of one row group: 8 adjacent outputs per lane from a register window; `src` / `dst` are complete shared
// addresses (ring buffer + row group + the lane's own offset) -----------------------------------------------------------------
template <int LXT, int LYT, int LZT>
__device__ __forceinline__ void s4_x_rows(const S3Params &P, const unsigned src, const unsigned dst, const int Lx) {
    typedef S3C<LXT, LYT, LZT> C;
    float v[C::WINX];
#pragma unroll
    for (int i = 0; i < C::WINX; i += 4) {
        const float4 t = s3_lds128(src + i * 4);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
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
    s3_sts128(dst, make_float4(a[0].x, a[0].y, a[1].x, a[1].y));
    s3_sts128(dst + 16, make_float4(a[2].x, a[2].y, a[3].x, a[3].y));
}

// One z step (see s3v_z_step) with the stores of v4: `op` points at row 0 of the EARLIER output plane; rows are P.row_b bytes
// apart, the later plane P.plane_b bytes further.  STEADY && full4: both planes are stored and all four rows are inside the
// volume (no predicates).
template <int LXT, int LYT, int LZT, bool CS, bool STEADY>
__device__ __forceinline__ void s4_z_step(const S3Params &P, const S3VTaps &T, float2 (&Pz)[4][S3VC<LXT, LYT, LZT>::NP],
                                          const float (&m0)[4], const float (&m1)[4], const int Lz, char *__restrict__ op,
                                          const int nrow, const bool full4, const bool emit0, const bool emit1) {
    typedef S3VC<LXT, LYT, LZT> C;
    constexpr int NP = C::NP;
    float2 fin[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        fin[v] = s3_fma2b(m0[v], T.p0[NP], Pz[v][NP - 1]);
        fin[v].x = fmaf(m1[v], T.p1[NP].x, fin[v].x);               // (kk[K-1], 0): the later output only
    }
#pragma unroll
    for (int i = NP - 1; i >= 1; --i) {
        if (LZT || 2 * i + 1 >= C::K - Lz) {
#pragma unroll
            for (int v = 0; v < 4; ++v) Pz[v][i] = s3_fma2b(m1[v], T.p1[i], s3_fma2b(m0[v], T.p0[i], Pz[v][i - 1]));
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        float2 z;
        z.x = 0.f;
        z.y = m0[v] * T.p0[0].y;                                    // (0, kk[0]) * m0
        Pz[v][0] = s3_fma2b(m1[v], T.p1[0], z);
    }
    auto put = [](char *q, float x) {
        if (CS) __stcs(reinterpret_cast<float *>(q), x); else *reinterpret_cast<float *>(q) = x;
    };
    if (STEADY && full4) {
        char *q = op, *r = op + P.plane_b;
        put(q, fin[0].y); put(r, fin[0].x);
#pragma unroll
        for (int v = 1; v < 4; ++v) {
            q += P.row_b; r += P.row_b;
            put(q, fin[v].y); put(r, fin[v].x);
        }
    } else {
        if (emit0) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (v < nrow) put(op + v * P.row_b, fin[v].y);
        }
        if (emit1) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (v < nrow) put(op + P.plane_b + v * P.row_b, fin[v].x);
        }
    }
}

template <int LXT, int LYT, int LZT, bool CS>
__global__ void __launch_bounds__(S3_NT, 1)
stream3d_kernel4(const __grid_constant__ S3Params P, const __grid_constant__ S3VTaps TZ, const __grid_constant__ CUtensorMap m_own,
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
    const int ncs = cl + (in_cols - cr), n1 = ncs * in_rows, wc = cr - cl;
    // Fast border patch (replicate / reflect / symmetric on 16-byte aligned tile edges): instead of walking a cell list, every
    // thread owns at most ONE patch item for the whole march, its offsets packed in three registers —
    //   X item: 4 adjacent out-of-range columns of one tile row <- 4 in-range cells (of the row's source row): 4 LDS.32 + 1 STS.128;
    //   Y item: 4 adjacent in-range columns of an out-of-range row <- the same columns of its source row:     1 LDS.128 + 1 STS.128.
    // Both kinds read in-range cells only and write out-of-range cells only, so they need no order among themselves, and ONLY
    // the threads that own an item wait for the planes' TMA barriers and issue the proxy fence: with the cell list every thread
    // of a border tile did both in every step, which (not the copies) was what made border tiles slower than interior ones —
    // and with 3.5 waves of tiles the slowest chain of tiles sets the launch time (1024^3 symmetric: 2.79 -> 2.62 ms; patching
    // before the arrival or by warp 0 alone were measured worse: 2.68 / 3.11 ms).
    unsigned fp_a = 0, fp_b = 0, fp_c = 0;       // dst | s0 << 16,  s1 | s2 << 16,  s3 | kind << 16
    bool fastfix = false;
    if (fix && tma && P.style != B2F_CIRCULAR) {
        const int gl = cl >> 2, gr = (in_cols - cr) >> 2, nxg = gl + gr;
        const int nro = rt + (in_rows - rb), nyg = wc >> 2;
        const int nx = in_rows * nxg, ny = nro * nyg;
        int bad = (((cl | cr | in_cols) & 3) != 0 || nx + ny > S3_NT) ? 1 : 0;
        if (!bad && tid < nx + ny) {
            int r, c0, kind;
            if (tid < nx) {
                r = tid / nxg;
                const int g = tid - r * nxg;
                c0 = g < gl ? 4 * g : cr + 4 * (g - gl);
                kind = 1;
            } else {
                const int t2 = tid - nx, rr = t2 / nyg;
                r = rr < rt ? rr : rb + (rr - rt);
                c0 = cl + 4 * (t2 - rr * nyg);
                kind = 2;
            }
            const int lr = (int)remap_index(P.style, (int64_t)ya + r, (int64_t)P.H) - ya;
            if (lr < rt || lr >= rb) bad = 1;
            int sc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int lc = c0 + k;
                if (kind == 1) {
                    lc = (int)remap_index(P.style, (int64_t)xa + c0 + k, (int64_t)P.W) - xa;
                    if (lc < cl || lc >= cr) bad = 1;
                }
                sc[k] = lr * S3_RWP + lc;
            }
            if (!bad) {
                fp_a = (unsigned)(r * S3_RWP + c0) | ((unsigned)sc[0] << 16);
                fp_b = (unsigned)sc[1] | ((unsigned)sc[2] << 16);
                fp_c = (unsigned)sc[3] | ((unsigned)kind << 16);
            }
        }
        fastfix = __syncthreads_or(bad) == 0;
    }
    const int ncell = (fix && !fastfix) ? n1 + (rt + (in_rows - rb)) * wc : 0;
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
        // source of the cell: >= 0 an offset inside the source plane in global memory, -1 the Fill value, <= -2 the cell
        // -(s + 2) of THIS raw tile: replicate / reflect / symmetric borders fold back into the in-range part of the tile, which
        // the TMA has delivered — the patch is then a shared-memory copy, no global load on the step's critical path
        int src = -1;
        if (sx >= 0 && sy >= 0) {
            const int lc = sx - xa, lr = sy - ya;
            src = (tma && lc >= cl && lc < cr && lr >= rt && lr < rb) ? -2 - (lr * S3_RWP + lc) : sy * P.W + sx;
        }
        cell_src[idx] = src;
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
            if (p >= 0 && p < in_planes) {
                s3_locate(P, zin0 + p, which, zz);
                // xy-filtered boundary planes (multi-GPU, see S3Params::xy_lo): 4 = a plane of xy_lo, 5 = a plane of xy_hi
                if (which == 1 && P.xy_lo) { which = 4; }
                else if (which == 2 && P.xy_hi) { which = 5; zz += P.xhi_o; }
                else if (which == 0 && P.xy_lo && zz < P.xlo_o) { which = 4; zz += P.xlo_h; }
                else if (which == 0 && P.xy_hi && zz >= P.own_n - P.xhi_o) { which = 5; zz -= P.own_n - P.xhi_o; }
            }
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
        if (which >= 4) {                       // xy-filtered plane: no TMA; a neighbour's plane may still be on its way
            const int zq = ptz[p & (S3_PT - 1)];
            if (which == 4 && zq < P.xlo_h && !lo_ready) { wait_flag(P.flag_lo + (y0 + TY > P.lo_early_rows ? 1 : 0)); lo_ready = true; }
            if (which == 5 && zq >= P.xhi_o && !hi_ready) { wait_flag(P.flag_hi); hi_ready = true; }
            s3_mbar_arrive(full + 8 * (p & (N - 1)));          // its ring slot still completes a phase: the parities stay in step
            return;
        }
        if (which == 1 && !lo_ready) { wait_flag(P.flag_lo + (ya + in_rows > P.lo_early_rows ? 1 : 0)); lo_ready = true; }
        if (which == 2 && !hi_ready) { wait_flag(P.flag_hi); hi_ready = true; }
        const int zz = which < 0 ? P.own_n : ptz[p & (S3_PT - 1)];         // Fill(0) plane: out of range reads zero
        const void *map = which == 1 ? (const void *)&m_lo : which == 2 ? (const void *)&m_hi : (const void *)&m_own;
        const int b = p & (N - 1);
        s3_mbar_expect_tx(full + 8 * b, (unsigned)C::RAWBYTES);
        s3_tma_load3d(raw_sa + b * (RAWSZ * 4), map, full + 8 * b, xa, ya, zz);
    };
    auto plane_src = [&](int p) -> const float * {
        const int which = ptw[p & (S3_PT - 1)];
        return which < 0 ? nullptr : (which == 1 ? P.lo : which == 2 ? P.hi : P.own) + (long long)ptz[p & (S3_PT - 1)] * P.plane;
    };
    // all threads: wait for the TMAs of input planes p .. p+n-1 (n = 1 or 2), then fill in their border cells from the gather
    // list; both planes share one pass over the list (all loads in flight together) and one proxy fence
    auto patch2 = [&](const int p, const int n) {
        if (fastfix) {                          // only the threads that own a patch item touch the planes (wait, copy, fence)
            if (fp_c >> 16) {
                s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
                if (n > 1) s3_mbar_wait(full + 8 * ((p + 1) & (N - 1)), ((p + 1) / N) & 1);
                float *d0 = raw + (p & (N - 1)) * RAWSZ, *d1 = raw + ((p + 1) & (N - 1)) * RAWSZ;
                const int dst = fp_a & 0xffff, s0 = fp_a >> 16;
                if ((fp_c >> 16) == 1) {
                    const int s1 = fp_b & 0xffff, s2 = fp_b >> 16, s3 = fp_c & 0xffff;
                    *reinterpret_cast<float4 *>(d0 + dst) = make_float4(d0[s0], d0[s1], d0[s2], d0[s3]);
                    if (n > 1) *reinterpret_cast<float4 *>(d1 + dst) = make_float4(d1[s0], d1[s1], d1[s2], d1[s3]);
                } else {
                    *reinterpret_cast<float4 *>(d0 + dst) = *reinterpret_cast<const float4 *>(d0 + s0);
                    if (n > 1) *reinterpret_cast<float4 *>(d1 + dst) = *reinterpret_cast<const float4 *>(d1 + s0);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            return;
        }
        if (tma) {
            s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1);
            if (n > 1) s3_mbar_wait(full + 8 * ((p + 1) & (N - 1)), ((p + 1) / N) & 1);
        }
        float *d0 = raw + (p & (N - 1)) * RAWSZ, *d1 = raw + ((p + 1) & (N - 1)) * RAWSZ;
        const float *g0 = plane_src(p), *g1 = n > 1 ? plane_src(p + 1) : nullptr;
#pragma unroll 2
        for (int idx = tid; idx < ncell; idx += S3_NT) {
            const int so = cell_src[idx], d = cell_dst[idx];
            float v0, v1 = 0.f;
            if (so <= -2) {
                v0 = d0[-2 - so];
                if (n > 1) v1 = d1[-2 - so];
            } else {
                v0 = (g0 != nullptr && so >= 0) ? __ldg(g0 + so) : P.fill;
                if (n > 1) v1 = (g1 != nullptr && so >= 0) ? __ldg(g1 + so) : P.fill;
            }
            d0[d] = v0;
            if (n > 1) d1[d] = v1;
        }
        if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };

    // prologue: planes 0..N-1 in flight.  Only thread 0 ever waits for a TMA to land: it checks the planes two steps before
    // stage x reads them and its next barrier arrival publishes that to the CTA (the other warps never touch the TMA
    // barriers: one try_wait latency less per warp-task).  Border tiles patch planes 0 .. 3 before the first stage x.
    auto landed = [&](int p) { s3_mbar_wait(full + 8 * (p & (N - 1)), (p / N) & 1); };
    const bool xy = P.xy_lo != nullptr || P.xy_hi != nullptr;
    auto is_pre = [&](int p) { return xy && ptw[p & (S3_PT - 1)] >= 4; };       // valid for the planes of the current table window
    if (tma && tid == 0) {
        for (int p = 0; p < min(N, in_planes); ++p) issue(p);
        for (int p = 0; p < min(4, in_planes); ++p)
            if (!is_pre(p)) landed(p);
    }
    auto patch_planes = [&](const int p, const int n) {             // like patch2, minus xy-filtered planes
        if (!xy) { patch2(p, n); return; }
        for (int k = 0; k < n; ++k)
            if (!is_pre(p + k)) patch2(p + k, 1);
    };
    if (fix) {                                  // planes 0 .. 3; step s patches planes 2s+6, 2s+7
        patch_planes(0, min(2, in_planes));
        if (in_planes > 2) patch_planes(2, min(2, in_planes - 2));
    }
    __syncthreads();

    // this thread's column and 4 rows in stages y / z
    const int gx = x0 + lane, gy = y0 + 4 * warp;
    const int nrow = (gx >= 0 && gx < P.W) ? min(4, P.H - gy) : 0;                  // <= 0: nothing to store
    // step s (planes 2s, 2s+1) completes the output planes o = 2s - (Lz-1) [by m0] and o + 1 [by m1], relative to zo0
    char *op = reinterpret_cast<char *>(P.out + ((long long)(zo0 - P.own_first) + (-2 - (Lz - 1))) * P.plane + (long long)gy * P.W + gx);   // step -1
    const bool full4 = nrow == 4;
    const int yoff = (4 * warp) * S3_XFP + lane;
    // stage x: the lane's offsets inside a row group of 8 rows (a quarter-warp covers two rows x 32 columns), kept in registers
    // through an opaque move (ptxas otherwise re-derives them from the thread index in every step)
    unsigned xsrc_l, xdst_l;
    {
        const int l8 = lane & 7, qw = lane >> 3, xg = l8 & 3, r = 2 * qw + (l8 >> 2);
        asm volatile("mov.u32 %0, %1;" : "=r"(xsrc_l) : "r"(raw_sa + (unsigned)(r * S3_RWP + 8 * xg) * 4u));
        asm volatile("mov.u32 %0, %1;" : "=r"(xdst_l) : "r"(xf_sa + (unsigned)(r * S3_XFP + 8 * xg) * 4u));
    }
    const int xrow_l = 2 * (lane >> 3) + ((lane & 7) >> 2);

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
        int t = warp - 1 - rot;
        if (t < 0) t += NXW;
        const int xn = bi == 2 ? 0 : 2 * bi + 2;                     // xf slots of planes pa, pa + 1
#pragma unroll 1
        for (; t < 2 * XW; t += NXW) {                               // task t: plane pa + (t >= XW), row group t mod XW
            const int d = t >= XW ? 1 : 0, g = t - d * XW;
            if ((d ? have_b : have_a) && ((LYT && C::RH % 8 == 0) || 8 * g + xrow_l < in_rows))
                s4_x_rows<LXT, LYT, LZT>(P, xsrc_l + (unsigned)(((pa + d) & (N - 1)) * (RAWSZ * 4) + g * (8 * S3_RWP * 4)),
                                         xdst_l + (unsigned)((xn + d) * (XFSZ * 4) + g * (8 * S3_XFP * 4)), Lx);
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
        op += 2 * P.plane_b;
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
                    if ((all || !xy) && (unsigned)rel < (unsigned)P.own_n) {     // an owned plane: no table, no flag
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
                if (all || (p < in_planes && !is_pre(p))) landed(p);
            }
        }
    };

    // ---- general step: every sub-step behind its run-time predicate (ramp-up, drain, volumes without TMA) --------------------
    auto slow_step = [&](const int s) {
        const int p0 = 2 * s;
        if (((p0 + 2) & (S3_PTB - 1)) == 0) locate_block(p0 + S3_PTA, S3_PTB);
        if (s >= 0 && !ok) s3_mbar_wait(xfull + 8 * bi, ph);
        ok = false;
        stage_x(p0 + 2, p0 + 2 < in_planes && !is_pre(p0 + 2), p0 + 3 < in_planes && !is_pre(p0 + 3));
        tma_check(p0, false);
        arrive_next();
        tma_work(p0, false);
        // border cells of the planes stage x reads TWO steps from now: behind the arrival, off the hand-off's critical path
        // (the next step's arrival publishes them)
        if (fix && p0 + 6 < in_planes) patch_planes(p0 + 6, min(2, in_planes - (p0 + 6)));
        if (s >= 0) {
            const int o = p0 - (Lz - 1);
            float2 ma[2], mb[2];
            const unsigned xa0 = xf_sa + (2 * bi * XFSZ + yoff) * 4;
            const bool pre_a = is_pre(p0), pre_b = p0 + 1 < in_planes && is_pre(p0 + 1);
            // an xy-filtered plane skips stages x and y: its values come straight from the xy buffer (L2, not L1: another
            // GPU's copy engine may have written them while this kernel was running)
            auto pre_load = [&](const int p, float2 (&m)[2]) {
                const float *b = (ptw[p & (S3_PT - 1)] == 4 ? P.xy_lo : P.xy_hi) + (long long)ptz[p & (S3_PT - 1)] * P.plane +
                                 (long long)gy * P.W + gx;
                float v[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) v[r] = r < nrow ? __ldcg(b + (long long)r * P.W) : 0.f;
                m[0] = make_float2(v[0], v[1]);
                m[1] = make_float2(v[2], v[3]);
            };
            if (!pre_a && p0 + 1 < in_planes && !pre_b) {
                s3v_y_task4x2<LXT, LYT, LZT>(P, xa0, xa0 + XFSZ * 4, ma, mb, Ly);
            } else {
                if (pre_a) pre_load(p0, ma); else s3_y_task4<LXT, LYT, LZT>(P, xa0, ma, Ly);
                if (p0 + 1 >= in_planes) mb[0] = mb[1] = make_float2(0.f, 0.f);
                else if (pre_b) pre_load(p0 + 1, mb);
                else s3_y_task4<LXT, LYT, LZT>(P, xa0 + XFSZ * 4, mb, Ly);
            }
            const float m0[4] = {ma[0].x, ma[0].y, ma[1].x, ma[1].y}, m1[4] = {mb[0].x, mb[0].y, mb[1].x, mb[1].y};
            // as in the steady state: the next step's barrier is tested a z stage ahead of its use (ramp-up and drain are 13 of
            // the 72 steps of a 128-plane slab)
            ok = s3_mbar_test(xfull + 8 * (bi == 2 ? 0 : bi + 1), bi == 2 ? (ph ^ 1) : ph);
            s4_z_step<LXT, LYT, LZT, CS, false>(P, TZ, acc, m0, m1, Lz, op, nrow, false, o >= 0 && o < nout && nrow > 0,
                                                o + 1 >= 0 && o + 1 < nout && nrow > 0);
        }
        advance();
    };
    // ---- steady state: both planes exist everywhere, both outputs are stored: no predicates ---------------------------------
    // `emit` = false: the ramp-up steps 0 .. Lz/2 - 1, whose outputs lie before the chunk (everything else as in the steady
    // state); `all` = false: the drain, where the planes 2s+2 .. 2s+N+1 may not exist any more (bounds checked).  The general
    // slow_step costs about twice a steady-state step, and ramp-up + drain are 13 of the 72 steps of a 128-plane slab.
    auto fast_step = [&](auto fixc, const int s, const bool emit, const bool all) {
        constexpr bool FIX = decltype(fixc)::value;
        const int p0 = 2 * s;
        if (!ok) s3_mbar_wait(xfull + 8 * bi, ph);
        stage_x(p0 + 2, all || p0 + 2 < in_planes, all || p0 + 3 < in_planes);
        tma_check(p0, all);
        arrive_next();
        tma_work(p0, all);
        if (FIX) {
            if (all) patch2(p0 + 6, 2);
            else if (p0 + 6 < in_planes) patch2(p0 + 6, min(2, in_planes - (p0 + 6)));
        }
        float2 ma[2], mb[2];
        const unsigned xa0 = xf_sa + (2 * bi * XFSZ + yoff) * 4;
        s3v_y_task4x2<LXT, LYT, LZT>(P, xa0, xa0 + XFSZ * 4, ma, mb, Ly);
        // the next step's barrier is tested here, a z stage ahead of its use: no warp sits out the barrier unit's latency
        ok = s3_mbar_test(xfull + 8 * (bi == 2 ? 0 : bi + 1), bi == 2 ? (ph ^ 1) : ph);
        const float m0[4] = {ma[0].x, ma[0].y, ma[1].x, ma[1].y}, m1[4] = {mb[0].x, mb[0].y, mb[1].x, mb[1].y};
        s4_z_step<LXT, LYT, LZT, CS, true>(P, TZ, acc, m0, m1, Lz, op, nrow, full4 && emit, emit && nrow > 0, emit && nrow > 0);
        advance();
    };

    // Steps s = -1 .. nsteps-1.  Steady state: s >= 0; both outputs exist, 0 <= 2s-(Lz-1) and 2s+1-(Lz-1) <= nout-1; the planes
    // up to 2s+N+1 exist.
    const int nsteps = (in_planes + 1) >> 1;
    int s_fast0 = nsteps, s_fast1 = nsteps;
    if (tma && nout + Lz - 3 >= 0 && in_planes - N - 2 >= 0) {
        s_fast0 = min(Lz >> 1, nsteps);                                      // ceil((Lz-1)/2)
        s_fast1 = max(s_fast0, min(min((nout + Lz - 3) / 2 + 1, (in_planes - N - 2) / 2 + 1), nsteps));
        if (xy) {                               // a steady-state step touches planes 2s .. 2s+N+1: all of them must be raw own planes
            const int ra = max(0, P.own_first + (P.xy_lo ? P.xlo_o : 0) - zin0);
            const int rb = min(in_planes, P.own_first + P.own_n - (P.xy_hi ? P.xhi_o : 0) - zin0);
            s_fast0 = min(max(s_fast0, (ra + 1) >> 1), nsteps);
            s_fast1 = max(s_fast0, min(s_fast1, (rb - N - 2) / 2 + 1));
            if (rb - N - 2 < 0) s_fast1 = s_fast0;
        }
    }
    // ramp-up on the steady-state code [s_pre0, s_fast0) and drain on it [s_fast1, s_drain): the steps whose two planes and two
    // outputs exist (xy-filtered boundary planes keep the general step)
    int s_pre0 = s_fast0, s_drain = s_fast1;
    if (s_fast1 > s_fast0) {
        s_pre0 = 0;
        s_drain = max(s_fast1, min((in_planes - 2) / 2, (nout + Lz - 3) / 2) + 1);
        if (xy) {
            const int ra = max(0, P.own_first + (P.xy_lo ? P.xlo_o : 0) - zin0);
            s_pre0 = min(s_fast0, (ra + 1) >> 1);
            s_drain = s_fast1;
        }
    }
    int s = -1;
    for (; s < s_pre0; ++s) slow_step(s);
    for (; s < s_fast0; ++s) {
        if (((2 * s + 2) & (S3_PTB - 1)) == 0) locate_block(2 * s + S3_PTA, S3_PTB);
        if (fix) fast_step(std::true_type{}, s, false, true); else fast_step(std::false_type{}, s, false, true);
    }
    while (s < s_fast1) {                       // the plane-source table is refilled every 8 steps, outside the inner loop
        if (((2 * s + 2) & (S3_PTB - 1)) == 0) locate_block(2 * s + S3_PTA, S3_PTB);
        const int e = min(s_fast1, (s + 1) | 7);
        if (fix) {
            for (; s < e; ++s) fast_step(std::true_type{}, s, true, true);
        } else {
            for (; s < e; ++s) fast_step(std::false_type{}, s, true, true);
        }
    }
    for (; s < s_drain; ++s) {
        if (((2 * s + 2) & (S3_PTB - 1)) == 0) locate_block(2 * s + S3_PTA, S3_PTB);
        if (fix) fast_step(std::true_type{}, s, true, false); else fast_step(std::false_type{}, s, true, false);
    }
    for (; s < nsteps; ++s) slow_step(s);
}

}  // namespace b2f
