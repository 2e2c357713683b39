// stream3d.cu — applicability, parameter set-up and launch of the fused 3-D separable kernel (stream3d.cuh)
#include <algorithm>
#include <cstdlib>

#include "stream3d_v4.cuh"

namespace b2f {


// Three 1-D factors on axes 0,1,2 in that order, <= 17 taps each, Float32 in and out, whole-array region.
bool stream3d_applicable(const Plan &P, int img_dt, int out_dt) {
    if (img_dt != B2F_F32 || out_dt != B2F_F32) return false;
    if (P.ndim != 3 || P.style > B2F_FILL || P.active.size() != 3) return false;
    for (int a = 0; a < 3; ++a) {
        const StageInfo &si = P.stages[P.active[a]];
        if (si.s->kind != B2F_STAGE_1D || si.s->axis != a) return false;
        if (si.s->len[a] < 1 || si.s->len[a] > S3_MAXTAPS) return false;
    }
    for (int d = 0; d < B2F_MAXDIM; ++d)
        if (P.roi.lo[d] != P.img_ax.lo[d] || P.roi.hi[d] != P.img_ax.hi[d] || P.out_ax.lo[d] != P.img_ax.lo[d] ||
            P.out_ax.hi[d] != P.img_ax.hi[d])
            return false;
    if (P.img_ax.len(0) >= (1LL << 30) || P.img_ax.len(1) >= (1LL << 30) || P.img_ax.len(2) >= (1LL << 30)) return false;
    if (P.img_ax.len(0) * P.img_ax.len(1) >= (1LL << 31)) return false;
    return true;
}

// the shape conditions of the TMA path, which the xy-filtered slab form needs (pointer alignment is the caller's: cudaMalloc'ed
// slabs qualify)
bool stream3d_xy_capable(const Plan &P) {
    if (!stream3d_applicable(P, B2F_F32, B2F_F32)) return false;
    const long long W = P.img_ax.len(0), H = P.img_ax.len(1);
    if (W % 4 != 0 || W < S3_RWP || H < S3_TY + S3_MAXTAPS - 1) return false;
    return P.style != B2F_FILL || (float)P.fill == 0.0f;
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*s3_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static s3_encode_fn s3_encoder() {
    static s3_encode_fn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (s3_encode_fn)f;
    }();
    return fn;
}

// 3-D map over `nplanes` dense W x H Float32 planes; box = (S3_RWP, rows, 1): one raw tile incl. halo, delivered with the
// shared-memory pitch the x stage expects; elements outside the array read as zero
static bool s3_make_map(CUtensorMap *m, const void *base, int W, int H, long long nplanes, int rows) {
    s3_encode_fn enc = s3_encoder();
    if (!enc || !base || nplanes < 1) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)S3_RWP, (cuuint32_t)rows, 1};
    if (W < S3_RWP || H < rows) return false;            // tiny arrays take the gather loader
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// z-chunks.  CTAs are dispatched in blockIdx order onto `slots` concurrent CTA slots.  A chunk re-runs Lz-1 planes of stages x
// and y, so long marches are cheapest, but whole waves of them quantise badly.  So: the first `nfull` tiles (whole waves)
// march all planes, the remaining tiles are cut into `kch` chunks each, which fills the last wave evenly.  (nfull, kch)
// minimise the makespan of that list schedule (closed form: w whole waves of full marches, then ceil(rest * k / slots) waves
// of chunks).
static int s3_geometry(S3Params &P, int ty_rows, int slots) {
    P.ntx = (P.W + P.xsh + S3_TX - 1) / S3_TX;
    P.nty = (P.H + ty_rows - 1) / ty_rows;
    const long long tiles = (long long)P.ntx * P.nty, own_n = P.own_n;
    const int ov = P.Lz - 1 + 3;
    long long best_full = tiles, best_k = 1, best_cost = ((tiles + slots - 1) / slots) * (own_n + ov);
    for (long long w = 0; w * slots <= tiles; ++w) {
        const long long rest = tiles - w * slots;
        if (rest == 0) break;
        for (long long k = 1; k <= 16 && k <= own_n; ++k) {
            const long long zc = (own_n + k - 1) / k;
            if (k > 1 && zc < 8) break;
            const long long c = w * (own_n + ov) + ((rest * k + slots - 1) / slots) * (zc + ov);
            if (c < best_cost) { best_cost = c; best_full = w * slots; best_k = k; }
        }
    }
    P.nfull = (int)best_full;
    P.kch = (int)best_k;
    P.zchunk = (int)((own_n + best_k - 1) / best_k);
    if (best_full + (tiles - best_full) * best_k > 0x7fffffffLL) return fail(B2F_ENOTSUP, "stream3d grid too large");
    return 0;
}

template <int LXT, int LYT, int LZT>
static int s3_launch_one(S3Params &P, const float *kz, cudaStream_t st) {
    typedef S3C<LXT, LYT, LZT> C;
    typedef S3VC<LXT, LYT, LZT> C4;
    // debugging knobs: B2F_S3_CS=0 plain instead of streaming stores, B2F_S3_NO_TMA forces the gather loader
    static const bool cs = getenv("B2F_S3_CS") ? atoi(getenv("B2F_S3_CS")) != 0 : true;
    static const bool no_tma = getenv("B2F_S3_NO_TMA") != nullptr;
    typedef void (*kern_t)(const S3Params, const S3VTaps, const CUtensorMap, const CUtensorMap, const CUtensorMap);
    const kern_t k4 = cs ? stream3d_kernel4<LXT, LYT, LZT, true> : stream3d_kernel4<LXT, LYT, LZT, false>;
    static thread_local bool configured = false;
    if (!configured) {
        B2F_CUDA(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C4::SMEM));
        configured = true;
    }
    S3VTaps TZ;
    memset(&TZ, 0, sizeof TZ);
    {   // z taps right-aligned in K (odd) slots, as the pairs of the paired z stage
        float kk[S3_MAXTAPS + 3] = {0};                      // kk[1 + j], j = -1 .. K
        for (int j = 0; j < P.Lz; ++j) kk[1 + C4::K - P.Lz + j] = kz[j];
        for (int i = 0; i <= C4::NP; ++i) {
            TZ.p0[i] = make_float2(kk[1 + 2 * i - 1], kk[1 + 2 * i]);
            TZ.p1[i] = make_float2(kk[1 + 2 * i], kk[1 + 2 * i + 1]);
        }
    }
    for (int j = 0; j < S3_MAXTAPS; ++j) P.kzr[j] = 0.f;
    for (int j = 0; j < P.Lz; ++j) P.kzr[C::LBZ - P.Lz + j] = kz[j];      // right-aligned in the LBZ slots
    alignas(64) CUtensorMap m_own, m_lo, m_hi;
    const bool tma_ok = !no_tma && P.use_tma && (P.style != B2F_FILL || P.fill == 0.0f);
    auto make_maps = [&](int rows) {
        memset(&m_own, 0, sizeof m_own); memset(&m_lo, 0, sizeof m_lo); memset(&m_hi, 0, sizeof m_hi);
        bool ok = tma_ok && s3_make_map(&m_own, P.own, P.W, P.H, P.own_n, rows);
        if (ok && P.lo_n > 0 && P.lo) ok = s3_make_map(&m_lo, P.lo, P.W, P.H, P.lo_n, rows);
        if (ok && P.hi_n > 0 && P.hi) ok = s3_make_map(&m_hi, P.hi, P.W, P.H, P.hi_n, rows);
        return ok;
    };
    // cp.async.bulk.tensor wants the box to start on a 16-byte boundary of the innermost axis (found the hard way:
    // "illegal instruction" otherwise), so with TMA the tile grid is shifted left by xsh = klox mod 4 columns
    const int xsh_tma = ((P.klox % 4) + 4) % 4, SMS = sm_count();
    const bool tma = make_maps(C::RH);
    P.use_tma = tma ? 1 : 0;
    if (!tma && (P.flag_lo || P.flag_hi || P.xy_lo || P.xy_hi))
        return fail(B2F_ENOTSUP, "staged / xy-filtered halos need the TMA path (row length a multiple of 4, 16-byte aligned buffers)");
    P.xsh = tma ? xsh_tma : 0;
    if (int rc = s3_geometry(P, S3_TY, SMS)) return rc;
    const long long nblocks = P.nfull + ((long long)P.ntx * P.nty - P.nfull) * P.kch;
    k4<<<(unsigned)nblocks, S3_NT, C4::SMEM, st>>>(P, TZ, m_own, m_lo, m_hi);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

// `own` holds planes [own_first, own_first+own_n) of a volume with Zg planes; lo/hi hold lo_n/hi_n planes below/above.
int run_stream3d_slab(const Plan &Pl, const void *own, const void *lo, int64_t lo_n, const void *hi, int64_t hi_n,
                      int64_t own_first, int64_t own_n, void *d_out, cudaStream_t st, const void *flag_lo, const void *flag_hi,
                      int epoch, int lo_early_rows, const b2f_slab_xy *xy) {
    S3Params P;
    memset(&P, 0, sizeof P);
    if (xy) {                                   // xy-filtered boundary planes: the halos exist as logical depths only
        P.xy_lo = (const float *)xy->xy_lo; P.xlo_h = (int)xy->lo_halo; P.xlo_o = (int)xy->lo_own;
        P.xy_hi = (const float *)xy->xy_hi; P.xhi_o = (int)xy->hi_own; P.xhi_h = (int)xy->hi_halo;
        lo = hi = nullptr;
        lo_n = xy->xy_lo ? xy->lo_halo : 0;
        hi_n = xy->xy_hi ? xy->hi_halo : 0;
    }
    P.flag_lo = lo_n > 0 ? (const unsigned char *)flag_lo : nullptr;
    P.flag_hi = hi_n > 0 ? (const unsigned char *)flag_hi : nullptr;
    P.epoch = epoch;
    P.lo_early_rows = lo_early_rows;
    P.own = (const float *)own; P.lo = (const float *)lo; P.hi = (const float *)hi;
    P.own_first = (int)own_first; P.own_n = (int)own_n; P.lo_n = (int)lo_n; P.hi_n = (int)hi_n;
    P.Zg = (int)Pl.img_ax.len(2);
    P.W = (int)Pl.img_ax.len(0); P.H = (int)Pl.img_ax.len(1);
    P.plane = (long long)P.W * P.H;
    P.row_b = (long long)P.W * 4;
    P.plane_b = P.plane * 4;
    P.out = (float *)d_out;
    P.style = Pl.style; P.fill = (float)Pl.fill;
    const StageInfo &sx = Pl.stages[Pl.active[0]], &sy = Pl.stages[Pl.active[1]], &sz = Pl.stages[Pl.active[2]];
    P.Lx = (int)sx.s->len[0]; P.klox = (int)sx.lo[0];
    P.Ly = (int)sy.s->len[1]; P.kloy = (int)sy.lo[1];
    P.Lz = (int)sz.s->len[2]; P.kloz = (int)sz.lo[2];
    float kz[S3_MAXTAPS];
    for (int j = 0; j < P.Lx; ++j) P.kx[j] = (float)sx.s->taps[j];
    for (int j = 1; j < P.Lx; ++j) P.kxp[j] = make_float2(P.kx[j], P.kx[j - 1]);
    for (int j = 0; j < P.Ly; ++j) P.ky[j] = (float)sy.s->taps[j];
    for (int j = 1; j < P.Ly; ++j) P.kyp[j] = make_float2(P.ky[j], P.ky[j - 1]);
    for (int j = 0; j < P.Lz; ++j) kz[j] = (float)sz.s->taps[j];
    auto al16 = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    P.use_tma = (P.W % 4 == 0) && al16(own) && (lo_n == 0 || !lo || al16(lo)) && (hi_n == 0 || !hi || al16(hi));
    P.vec_out = (P.W % 2 == 0) && reinterpret_cast<uintptr_t>(d_out) % 8 == 0;
    if (P.Lx == 17 && P.Ly == 17 && P.Lz == 17) return s3_launch_one<17, 17, 17>(P, kz, st);
    if (P.Lx == 9 && P.Ly == 9 && P.Lz == 9) return s3_launch_one<9, 9, 9>(P, kz, st);
    if (P.Lx == 5 && P.Ly == 5 && P.Lz == 5) return s3_launch_one<5, 5, 5>(P, kz, st);
    if (P.Lx == 3 && P.Ly == 3 && P.Lz == 3) return s3_launch_one<3, 3, 3>(P, kz, st);
    return s3_launch_one<0, 0, 0>(P, kz, st);
}

int run_stream3d(const Plan &P, const void *d_img, void *d_out, cudaStream_t st) {
    set_path("stream3d");
    return run_stream3d_slab(P, d_img, nullptr, 0, nullptr, 0, 0, P.img_ax.len(2), d_out, st);
}

}  // namespace b2f
