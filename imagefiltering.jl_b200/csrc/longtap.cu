// longtap.cu — one 1-D stage of 18 .. 256 taps along any axis of an N-d array (K1-long).
//
// The reference runs a separable cascade one 1-D factor at a time whatever the factor's length (src/imfilter.jl:385-395,
// inner loop :724-739); its own benchmark kernel KernelFactors.gaussian(sigma = 10) is 41 taps per axis
// (benchmark/benchmarks.jl:45-49).  The register-window kernels (stream2d / stream3d) unroll over the tap count and stop at
// 17; this kernel keeps the OUTPUTS in registers and walks the taps in chunks instead, so its code size is independent of
// the tap count:
//   * the array is viewed as (W, H, B) with the filtered axis = H (axis >= 1: W = product of the leading extents) or as rows
//     of the contiguous axis (axis 0);
//   * a CTA (256 threads) owns 32 lines x 128 outputs along the axis.  The input tile (128 + L - 1 positions, border remap and
//     eltype conversion applied once, at load) sits in shared memory as [position][line] with a pitch of 33, so that a warp —
//     lane = line — reads one position of 32 lines without bank conflicts, whichever axis is filtered; along axis 0 the
//     result goes back through the same transposed layout for coalesced stores;
//   * a thread owns 8 consecutive outputs of one line.  Float32: four float2 accumulators fed by FFMA2 in the
//     value-broadcast x tap-pair form (value u = in[o + j] serves outputs o and o+1 with taps (k[j], k[j-1])), a sliding
//     register window of 14 values per chunk of 8 taps: 32 FFMA2 per 8 shared-memory loads.  Float64: eight accumulators,
//     separate multiply and add in tap order (the reference's arithmetic, src/imfilter.jl:732-737), window of 16 per 8 taps.
// Roofline: one pass moves sizeof(in) + sizeof(out) bytes per element and issues L multiply-adds; at 41 taps Float32 the FP32
// pipe (128 FMA / clk / SM) and HBM are within 15 % of each other.
#include <type_traits>

#include "common.cuh"

namespace b2f {

constexpr int LT_MAXL = 256, LT_MINL = 18;
constexpr int LT_NT = 256, LT_TO = 128, LT_PITCH = 33;     // (LT_TO = 256 gains 4 % on 8192^2 and loses 40 % on 100^3)
constexpr int LT_NG = LT_TO / 64;            // groups of 8 outputs per warp

template <typename CT>
struct LtParams {
    const void *src;
    CT *dst;
    int src_dt, L, klo, style, along_x;
    long long W, H, B;          // the (W, H, B) view; along_x: W = the axis, H = number of rows, B = 1
    long long Ag, a_first;      // global length of the filtered axis and the global index of local position 0 (slab form)
    long long o0, on;           // outputs [o0, o0 + on) along the axis (local positions); the output array has `on` of them
    long long ntl, nta;         // tiles across the lines / along the axis
    CT fill;
    CT k[LT_MAXL];
};

__device__ __forceinline__ float2 lt_fma2(float v, float2 k, float2 c) {
    unsigned long long rk = *reinterpret_cast<unsigned long long *>(&k), rc = *reinterpret_cast<unsigned long long *>(&c), rv, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rv), "l"(rk), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// 8 outputs of one line: sp = the line's tile entry of the first output's first input, positions LT_PITCH apart
__device__ __forceinline__ void lt_line8(const float *sp, const float *k, const float2 *kp, const int L, float (&out)[8]) {
    float2 acc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[p] = make_float2(sp[(2 * p) * LT_PITCH] * k[0], 0.f);      // j = 0: the even output only
    float w[14];
#pragma unroll
    for (int t = 0; t < 6; ++t) w[t] = sp[(1 + t) * LT_PITCH];
    int j0 = 1;
    for (; j0 + 8 <= L; j0 += 8) {                                   // taps j0 .. j0+7 (all < L): value u = 2p + j is w[u - j0]
#pragma unroll
        for (int t = 0; t < 8; ++t) w[6 + t] = sp[(j0 + 6 + t) * LT_PITCH];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const float2 kk = kp[j0 + jj];
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[p] = lt_fma2(w[2 * p + jj], kk, acc[p]);
        }
#pragma unroll
        for (int t = 0; t < 6; ++t) w[t] = w[8 + t];
    }
    if (j0 < L) {                                                    // the last, partial chunk (the tile is padded: loads stay inside)
#pragma unroll
        for (int t = 0; t < 8; ++t) w[6 + t] = sp[(j0 + 6 + t) * LT_PITCH];
#pragma unroll
        for (int jj = 0; jj < 7; ++jj) {
            if (j0 + jj < L) {
                const float2 kk = kp[j0 + jj];
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = lt_fma2(w[2 * p + jj], kk, acc[p]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {                                    // j = L: the odd output only
        acc[p].y = fmaf(sp[(2 * p + L) * LT_PITCH], k[L - 1], acc[p].y);
        out[2 * p] = acc[p].x;
        out[2 * p + 1] = acc[p].y;
    }
}

__device__ __forceinline__ void lt_line8(const double *sp, const double *k, const double *, const int L, double (&out)[8]) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
    double w[16];
#pragma unroll
    for (int t = 0; t < 8; ++t) w[t] = sp[t * LT_PITCH];
    int j0 = 0;
    for (; j0 + 8 <= L; j0 += 8) {
#pragma unroll
        for (int t = 0; t < 8; ++t) w[8 + t] = sp[(j0 + 8 + t) * LT_PITCH];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const double kk = k[j0 + jj];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = mac<double>(acc[i], w[i + jj], kk);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) w[t] = w[8 + t];
    }
    if (j0 < L) {
#pragma unroll
        for (int t = 0; t < 8; ++t) w[8 + t] = sp[(j0 + 8 + t) * LT_PITCH];
#pragma unroll
        for (int jj = 0; jj < 7; ++jj) {
            if (j0 + jj < L) {
                const double kk = k[j0 + jj];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = mac<double>(acc[i], w[i + jj], kk);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = acc[i];
}

template <typename CT> struct LtPair { typedef CT type; };
template <> struct LtPair<float> { typedef float2 type; };

// Border remap in 32-bit arithmetic (axes are < 2^31, checked by the host side): the in-range test inline, the folds out of line
// (common.cuh's 64-bit remap_index inlines two 64-bit divisions into every unrolled load).
static __device__ __noinline__ int lt_remap_slow(int style, int i, int n) { return (int)remap_index(style, (int64_t)i, (int64_t)n); }
__device__ __forceinline__ int lt_remap(int style, int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    return lt_remap_slow(style, i, n);
}

// The tile load.  Every thread issues its loads in batches of independent, unconditional loads (cells outside the array read
// element 0 of their line and are replaced afterwards): a dependent address -> load -> store chain per cell would cost a full
// DRAM round trip each.  Cell codes: >= 0 source position, -1 Fill value, -2 not a cell of this thread, -3 padding (zero).
template <typename IT, bool N0, typename CT>
__device__ __forceinline__ void lt_load_tile(const LtParams<CT> &P, CT *tile, const long long line0, const int out0, const long long b,
                                             const long long nlines, const int npos, const int npos_pad) {
    const IT *src = reinterpret_cast<const IT *>(P.src);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto conv = [](IT x) -> CT { return N0 ? N0f8Conv<CT>::f((unsigned)x) : (CT)x; };
    // source eltype == compute type (Float32 -> Float32, Float64 -> Float64; every pass after the first): the cells go from global
    // to shared memory with cp.async — all of a thread's ~24 cells in flight at once, no staging registers, one wait
    if (std::is_same<IT, CT>::value && !N0) {
        auto cp_cell = [](CT *dst, const IT *g) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            if (sizeof(CT) == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
        };
        // interior tiles (no border, all 32 lines present): straight pointer walks, no remap, no predicates per cell
        const long long axis_len = P.along_x ? P.W : P.Ag;
        const long long first_pos = (P.along_x ? 0 : P.a_first) + out0 + P.klo;
        const bool interior = first_pos >= 0 && first_pos + npos <= axis_len && line0 + 32 <= nlines &&
                              (P.along_x || (first_pos - P.a_first >= 0 && first_pos - P.a_first + npos <= P.H));
        if (interior && P.along_x) {
            for (int ln = warp; ln < 32; ln += LT_NT / 32) {
                const IT *g = src + (line0 + ln) * P.W + first_pos + lane;
                CT *d = tile + lane * LT_PITCH + ln;
                int q = lane;
                for (; q < npos; q += 32, g += 32, d += 32 * LT_PITCH) cp_cell(d, g);
                for (; q < npos_pad; q += 32, d += 32 * LT_PITCH) *d = (CT)0;
            }
        } else if (interior) {
            const IT *g = src + (b * P.H + (first_pos - P.a_first) + warp) * P.W + line0 + lane;
            CT *d = tile + warp * LT_PITCH + lane;
            const long long gs = (long long)(LT_NT / 32) * P.W;
            int q = warp;
            for (; q < npos; q += LT_NT / 32, g += gs, d += (LT_NT / 32) * LT_PITCH) cp_cell(d, g);
            for (; q < npos_pad; q += LT_NT / 32, d += (LT_NT / 32) * LT_PITCH) *d = (CT)0;
        } else if (P.along_x) {
            const int Wi = (int)P.W, p0 = out0 + P.klo;
            for (int ln = warp; ln < 32; ln += LT_NT / 32) {
                const long long row = line0 + ln;
                const bool rowok = row < nlines;
                const IT *rp = src + (rowok ? row * P.W : 0);
                for (int q = lane; q < npos_pad; q += 32) {
                    int a = -3;
                    if (q < npos && rowok) a = lt_remap(P.style, p0 + q, Wi);
                    CT *d = tile + q * LT_PITCH + ln;
                    if (a >= 0) cp_cell(d, rp + a); else *d = a == -1 ? P.fill : (CT)0;
                }
            }
        } else {
            const long long xg = line0 + lane;
            const bool xin = xg < nlines;
            const int Hi = (int)P.H, Agi = (int)P.Ag, af = (int)P.a_first, p0 = af + out0 + P.klo;
            const IT *cp = src + b * P.H * P.W + (xin ? xg : 0);
            for (int q = warp; q < npos_pad; q += LT_NT / 32) {
                int a = -3;
                if (q < npos && xin) {
                    int r = lt_remap(P.style, p0 + q, Agi);
                    if (r >= 0) r -= af;
                    a = (r >= 0 && r < Hi) ? r : -1;
                }
                CT *d = tile + q * LT_PITCH + lane;
                if (a >= 0) cp_cell(d, cp + (long long)a * P.W); else *d = a == -1 ? P.fill : (CT)0;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }
    if (P.along_x) {
        // lines = rows (stride W), positions contiguous: lanes run along the positions (coalesced), smem [pos][line]
        constexpr int UB = 6;
        const int Wi = (int)P.W, p0 = out0 + P.klo;
#pragma unroll 1
        for (int ln = warp; ln < 32; ln += LT_NT / 32) {
            const long long row = line0 + ln;
            const bool rowok = row < nlines;
            const IT *rp = src + (rowok ? row * P.W : 0);
#pragma unroll 1
            for (int q0 = lane; q0 < npos_pad; q0 += 32 * UB) {
                IT x[UB];
                int a[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int q = q0 + 32 * u;
                    a[u] = q < npos_pad ? -3 : -2;
                    if (q < npos && rowok) a[u] = lt_remap(P.style, p0 + q, Wi);
                    x[u] = rp[a[u] >= 0 ? a[u] : 0];
                }
#pragma unroll
                for (int u = 0; u < UB; ++u)
                    if (a[u] != -2) tile[(q0 + 32 * u) * LT_PITCH + ln] = a[u] >= 0 ? conv(x[u]) : (a[u] == -1 ? P.fill : (CT)0);
            }
        }
    } else {
        constexpr int UB = 8, NW = LT_NT / 32;
        const long long xg = line0 + lane;
        const bool xin = xg < nlines;
        const int Hi = (int)P.H, Agi = (int)P.Ag, af = (int)P.a_first, p0 = af + out0 + P.klo;
        const IT *cp = src + b * P.H * P.W + (xin ? xg : 0);
#pragma unroll 1
        for (int q0 = warp; q0 < npos_pad; q0 += NW * UB) {
            IT x[UB];
            int a[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int q = q0 + NW * u;
                a[u] = q < npos_pad ? -3 : -2;
                if (q < npos && xin) {
                    int r = lt_remap(P.style, p0 + q, Agi);
                    if (r >= 0) r -= af;
                    // a slab holds the planes its outputs need (checked by the caller); anything else is a Fill cell
                    a[u] = (r >= 0 && r < Hi) ? r : -1;
                }
                x[u] = cp[(long long)(a[u] >= 0 ? a[u] : 0) * P.W];
            }
#pragma unroll
            for (int u = 0; u < UB; ++u)
                if (a[u] != -2) tile[(q0 + NW * u) * LT_PITCH + lane] = a[u] >= 0 ? conv(x[u]) : (a[u] == -1 ? P.fill : (CT)0);
        }
    }
}

// shared memory: taps k[Lp] | (Float32) tap pairs kp[Lp] | tile [npos_pad][33]
template <typename CT>
__global__ void __launch_bounds__(LT_NT) longtap_kernel(const __grid_constant__ LtParams<CT> P) {
    typedef typename LtPair<CT>::type PT;
    extern __shared__ __align__(16) unsigned char lt_smem[];
    const int L = P.L, Lp = (L + 8) & ~7;
    CT *k = reinterpret_cast<CT *>(lt_smem);
    PT *kp = reinterpret_cast<PT *>(k + Lp);
    CT *tile = reinterpret_cast<CT *>(kp + (sizeof(CT) == 4 ? Lp : 0));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int npos = LT_TO + L - 1, npos_pad = LT_TO + Lp + 8;

    long long t = blockIdx.x;
    const long long ta = t % P.nta; t /= P.nta;
    const long long tl = t % P.ntl;
    const long long b = t / P.ntl;
    const long long line0 = tl * 32;                                  // first line of this tile
    const int out0 = (int)(P.o0 + ta * LT_TO);                        // ... and its first output position (axes are < 2^31)
    const long long nlines = P.along_x ? P.H : P.W;

    for (int j = tid; j < Lp; j += LT_NT) {
        k[j] = j < L ? P.k[j] : (CT)0;
        if (sizeof(CT) == 4) {
            float2 pr = make_float2(j < L ? (float)P.k[j] : 0.f, (j >= 1 && j <= L) ? (float)P.k[j - 1] : 0.f);
            reinterpret_cast<float2 *>(kp)[j] = pr;
        }
    }
    // ---- tile load: position q of the tile is input position out0 + klo + q ------------------------------------------------
    switch (P.src_dt) {
        case B2F_U8: lt_load_tile<uint8_t, false>(P, tile, line0, out0, b, nlines, npos, npos_pad); break;
        case B2F_N0F8: lt_load_tile<uint8_t, true>(P, tile, line0, out0, b, nlines, npos, npos_pad); break;
        case B2F_F32: lt_load_tile<float, false>(P, tile, line0, out0, b, nlines, npos, npos_pad); break;
        default: lt_load_tile<double, false>(P, tile, line0, out0, b, nlines, npos, npos_pad); break;
    }
    __syncthreads();
    // ---- compute: group g = 8 consecutive outputs; LT_TO / 8 groups per tile, LT_NG per warp ------------------------------------
    const int oend = (int)(P.o0 + P.on);
    if (!P.along_x) {                               // lane = x: one coalesced row per output
        const long long x = line0 + lane;
        CT res[LT_NG][8];
#pragma unroll
        for (int h = 0; h < LT_NG; ++h) lt_line8(tile + (8 * (warp + 8 * h)) * LT_PITCH + lane, k, kp, L, res[h]);
        if (x < nlines) {
#pragma unroll
            for (int h = 0; h < LT_NG; ++h) {
                const int o = out0 + 8 * (warp + 8 * h);
                CT *dp = P.dst + ((b * P.on + (o - (int)P.o0)) * P.W + x);
#pragma unroll
                for (int i = 0; i < 8; ++i, dp += P.W)
                    if (o + i < oend) *dp = res[h][i];
            }
        }
    } else {                                        // lane = row: the results go back through the transposed tile
        CT res[LT_NG][8];
#pragma unroll
        for (int h = 0; h < LT_NG; ++h) lt_line8(tile + (8 * (warp + 8 * h)) * LT_PITCH + lane, k, kp, L, res[h]);
        __syncthreads();                            // every warp is done reading the tile: reuse its first LT_TO positions
#pragma unroll
        for (int h = 0; h < LT_NG; ++h) {
            const int g = warp + 8 * h;
#pragma unroll
            for (int i = 0; i < 8; ++i) tile[(8 * g + i) * LT_PITCH + lane] = res[h][i];
        }
        __syncthreads();
        for (int idx = tid; idx < 32 * LT_TO; idx += LT_NT) {   // lanes along the positions: coalesced stores
            const int ln = idx / LT_TO, q = idx - ln * LT_TO;
            const long long row = line0 + ln;
            if (row < nlines && out0 + q < oend) P.dst[row * P.on + (out0 + q - (int)P.o0)] = tile[q * LT_PITCH + ln];
        }
    }
}

bool longtap_ok(int64_t L) { return L >= LT_MINL && L <= LT_MAXL; }

// One pass: `taps` (L of them, first tap at offset klo) along the H axis of the (W, H, B) view, or along W when along_x.
// Slab form (axis >= 1 only): the source holds H positions starting at global position a_first of an axis of global length
// Ag; outputs [o0, o0 + on) in local positions.  Ag = 0: the axis is whole (Ag = H, a_first = 0, all outputs).
template <typename CT>
int run_longtap(const void *src, int src_dt, CT *dst, const double *taps, int64_t L, int64_t klo, bool along_x, int64_t W, int64_t H,
                int64_t B, int style, CT fill, int64_t Ag, int64_t a_first, int64_t o0, int64_t on, cudaStream_t st) {
    if (!longtap_ok(L)) return fail(B2F_ENOTSUP, "longtap: %lld taps (takes %d .. %d)", (long long)L, LT_MINL, LT_MAXL);
    LtParams<CT> P;
    memset(&P, 0, sizeof P);
    P.src = src; P.dst = dst; P.src_dt = src_dt; P.L = (int)L; P.klo = (int)klo; P.style = style; P.along_x = along_x ? 1 : 0;
    P.W = W; P.H = H; P.B = along_x ? 1 : B;
    const int64_t axis_len = along_x ? W : H;
    P.Ag = Ag > 0 ? Ag : axis_len; P.a_first = Ag > 0 ? a_first : 0;
    P.o0 = Ag > 0 ? o0 : 0; P.on = Ag > 0 ? on : axis_len;
    P.fill = fill;
    for (int j = 0; j < L; ++j) P.k[j] = (CT)taps[j];
    const int64_t nlines = along_x ? H : W;
    P.ntl = (nlines + 31) / 32;
    P.nta = (P.on + LT_TO - 1) / LT_TO;
    const long long blocks = P.ntl * P.nta * P.B;
    if (blocks <= 0) return 0;
    if (blocks >= (1LL << 31) || axis_len >= (1LL << 31) - 4096 || P.Ag >= (1LL << 31) - 4096)
        return fail(B2F_ENOTSUP, "longtap: array too large for one launch");
    const int Lp = ((int)L + 8) & ~7;
    const size_t smem = (size_t)Lp * sizeof(CT) * (sizeof(CT) == 4 ? 3 : 1) + (size_t)(LT_TO + Lp + 8) * LT_PITCH * sizeof(CT);
    if (smem > 48 * 1024)                        // per launch: the attribute belongs to the current device's copy of the kernel
        B2F_CUDA(cudaFuncSetAttribute(longtap_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    longtap_kernel<CT><<<(unsigned)blocks, LT_NT, smem, st>>>(P);
    count_launch(1);
    B2F_CUDA(cudaGetLastError());
    return 0;
}

template int run_longtap<float>(const void *, int, float *, const double *, int64_t, int64_t, bool, int64_t, int64_t, int64_t, int, float,
                                int64_t, int64_t, int64_t, int64_t, cudaStream_t);
template int run_longtap<double>(const void *, int, double *, const double *, int64_t, int64_t, bool, int64_t, int64_t, int64_t, int,
                                 double, int64_t, int64_t, int64_t, int64_t, cudaStream_t);

}  // namespace b2f
