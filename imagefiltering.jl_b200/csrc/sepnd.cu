// sepnd.cu — N-d separable cascades as a chain of streamed passes, and the slab (multi-GPU) form.
//
// A cascade of 1-D factors on distinct axes (KernelFactors.gaussian((4,4,4)) and friends; reference scheduler
// src/imfilter.jl:385-395,438-446) is run as: [fused x+y pass if the cascade starts with axes 0,1] then one
// streamed pass per remaining stage, each reading the previous pass's output (eltype(out) temporaries, like the
// reference's `tempbuffer`, src/imfilter.jl:1317-1329).  A stage along axis a >= 1 sees the array as a 2-D image
// of width prod(dims[0..a-1]) whose "rows" are the slices along a, so it is the y-only mode of stream2d.cuh.
//
// Border handling per pass instead of "pad once": identical results because every axis is filtered by at most one
// stage (remaps are per-axis and commute with filtering along other axes).  For Fill(v) the
// out-of-range value a later stage must see is v pushed through the earlier stages, i.e. sum_j v*k[j] accumulated
// in tap order — computed here on the host with the same arithmetic as the device.
//
// b2f_imfilter_slab is the same y-only pass over a buffer that holds a rank's planes plus halo planes received
// from its neighbours; the border along the sharded axis is evaluated in GLOBAL plane coordinates.
#include <cmath>

#include "stream2d.cuh"

namespace b2f {

template <typename IT, typename CT, int NPL> int launch_stream2d(const S2Params<CT, NPL> &P, cudaStream_t st);
template <typename IT, typename CT> int launch_stream1d(const S2Params<CT, 1> &P, bool along_x, cudaStream_t st);
#define B2F_S2_DECL(IT, CT)                                                                  \
    template <> int launch_stream2d<IT, CT, 1>(const S2Params<CT, 1> &, cudaStream_t);       \
    template <> int launch_stream1d<IT, CT>(const S2Params<CT, 1> &, bool, cudaStream_t);
B2F_S2_DECL(uint8_t, float)
B2F_S2_DECL(uint8_t, double)
B2F_S2_DECL(float, float)
B2F_S2_DECL(float, double)
B2F_S2_DECL(double, float)
B2F_S2_DECL(double, double)

// 18 .. 256 taps: the chunked-tap kernel of longtap.cu (one pass per stage)
bool longtap_ok(int64_t L);
template <typename CT>
int run_longtap(const void *src, int src_dt, CT *dst, const double *taps, int64_t L, int64_t klo, bool along_x, int64_t W, int64_t H,
                int64_t B, int style, CT fill, int64_t Ag, int64_t a_first, int64_t o0, int64_t on, cudaStream_t st);

static bool taps_ok_single(int64_t L) { return L >= 1 && (L <= 17 || longtap_ok(L)); }
static bool taps_ok_pair(int64_t Lx, int64_t Ly) { return (Lx <= 16 && Ly <= 16) || (Lx == 17 && Ly == 17); }

bool sepnd_applicable(const Plan &P, int img_dt, int out_dt) {
    if (out_dt != B2F_F32 && out_dt != B2F_F64) return false;
    if (img_dt != B2F_U8 && img_dt != B2F_N0F8 && img_dt != B2F_F32 && img_dt != B2F_F64) return false;
    if (P.style > B2F_FILL || P.active.empty()) return false;
    bool seen[B2F_MAXDIM] = {false, false, false, false};
    for (int a : P.active) {
        const StageInfo &si = P.stages[a];
        if (si.s->kind != B2F_STAGE_1D) return false;
        const int ax = si.s->axis;
        if (seen[ax] || !taps_ok_single(si.s->len[ax])) return false;
        seen[ax] = true;
    }
    for (int d = 0; d < B2F_MAXDIM; ++d)
        if (P.roi.lo[d] != P.img_ax.lo[d] || P.roi.hi[d] != P.img_ax.hi[d] || P.out_ax.lo[d] != P.img_ax.lo[d] ||
            P.out_ax.hi[d] != P.img_ax.hi[d])
            return false;
    if (P.img_ax.count() >= (1LL << 40)) return false;
    return true;
}

template <typename CT> static CT push_fill(CT v, const double *taps, int64_t L);
template <> double push_fill<double>(double v, const double *taps, int64_t L) {
    volatile double acc = 0.0;
    for (int64_t j = 0; j < L; ++j) { volatile double p = v * taps[j]; acc = acc + p; }
    return acc;
}
template <> float push_fill<float>(float v, const double *taps, int64_t L) {
    float acc = 0.0f;
    for (int64_t j = 0; j < L; ++j) acc = std::fmaf(v, (float)taps[j], acc);
    return acc;
}

template <typename CT>
static void strip_geometry(S2Params<CT, 1> &P, long long nbatch) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int CW = 32 * PX;
    P.nsx = (P.rw + CW - 1) / CW;
    const long long want = (long long)sm_count() * 16 * 6;
    int SH = 256;
    while (SH > 32 && (long long)P.nsx * ((P.rh + SH - 1) / SH) * nbatch < want) SH >>= 1;
    P.SH = SH;
    P.nsy = (P.rh + SH - 1) / SH;
    P.nstrips = (long long)P.nsx * P.nsy * nbatch;
}

template <typename CT>
static int launch_by_input(const S2Params<CT, 1> &P, int src_dt, int mode /*0 xy, 1 x, 2 y*/, cudaStream_t st) {
    switch (src_dt) {
        case B2F_U8: case B2F_N0F8:
            return mode == 0 ? launch_stream2d<uint8_t, CT, 1>(P, st) : launch_stream1d<uint8_t, CT>(P, mode == 1, st);
        case B2F_F32:
            return mode == 0 ? launch_stream2d<float, CT, 1>(P, st) : launch_stream1d<float, CT>(P, mode == 1, st);
        default:
            return mode == 0 ? launch_stream2d<double, CT, 1>(P, st) : launch_stream1d<double, CT>(P, mode == 1, st);
    }
}

// one pass: stage sx along x (or null) and stage sy along the axis `yaxis` (or null) over a dense array `dims`
template <typename CT>
static int run_pass(const int64_t *dims, const StageInfo *sx, const StageInfo *sy, int yaxis, const void *src, int src_dt,
                    void *dst, int style, CT fill, int64_t Hg, int64_t y_first, int64_t ry0, int64_t rh, cudaStream_t st) {
    constexpr int PX = S2Vec<CT>::PX;
    S2Params<CT, 1> P;
    memset(&P, 0, sizeof P);
    long long W, H, nbatch;
    if (sy) {
        W = 1; for (int d = 0; d < yaxis; ++d) W *= dims[d];
        H = dims[yaxis];
        nbatch = 1; for (int d = yaxis + 1; d < B2F_MAXDIM; ++d) nbatch *= dims[d];
    } else {
        W = dims[0]; H = dims[1] * dims[2] * dims[3]; nbatch = 1;
    }
    if (sx && !sy && sx->s->len[0] > 17)
        return run_longtap<CT>(src, src_dt, (CT *)dst, sx->s->taps, sx->s->len[0], sx->lo[0], true, W, H, 1, style, fill, 0, 0, 0, 0, st);
    if (sy && !sx && sy->s->len[yaxis] > 17)
        return run_longtap<CT>(src, src_dt, (CT *)dst, sy->s->taps, sy->s->len[yaxis], sy->lo[yaxis], false, W, H, nbatch, style, fill,
                               Hg > 0 ? Hg : 0, y_first, ry0, rh > 0 ? rh : H, st);
    if (W >= (1LL << 30) || H >= (1LL << 30)) return fail(B2F_ENOTSUP, "array extent too large for the streamed pass");
    P.img = src;
    P.n0_r = src_dt == B2F_N0F8 ? (CT)1 / (CT)255 : (CT)1;
    P.n0_c = src_dt == B2F_N0F8 ? (CT)255 : (CT)1;
    P.W = (int)W; P.H = (int)H;
    P.Hg = (int)(Hg > 0 ? Hg : H); P.y_first = (int)y_first;
    P.img_plane = W * H;
    P.out[0] = dst;
    P.rx0 = 0; P.rw = (int)W;
    P.ry0 = (int)ry0; P.rh = (int)(rh > 0 ? rh : H);
    P.out_ox = 0; P.out_oy = (int)ry0;
    P.out_pitch = W; P.out_plane = W * P.rh;
    P.style = style; P.fill = fill;
    P.fma = accum_mode() == B2F_ACCUM_FMA;      // honoured by the fused x+y pass for the instantiated tap counts (b2f_set_accum_mode)
    P.Lx = 1; P.Ly = 1;
    if (sx) {
        P.Lx = (int)sx->s->len[0]; P.klox = (int)sx->lo[0];
        for (int j = 0; j < P.Lx; ++j) P.kx[0][j] = (CT)sx->s->taps[j];
        for (int j = 1; j < P.Lx; ++j) P.kxp[0][j] = make_float2((float)sx->s->taps[j], (float)sx->s->taps[j - 1]);
    }
    if (sy) {
        P.Ly = (int)sy->s->len[yaxis]; P.kloy = (int)sy->lo[yaxis];
        for (int j = 0; j < P.Ly; ++j) P.ky[0][j] = (CT)sy->s->taps[j];
    }
    P.vec_ok = (W % PX == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
    strip_geometry(P, nbatch);
    return launch_by_input<CT>(P, src_dt, sx && sy ? 0 : (sx ? 1 : 2), st);
}

template <typename CT> struct CtDt;
template <> struct CtDt<float> { static const int v = B2F_F32; };
template <> struct CtDt<double> { static const int v = B2F_F64; };

template <typename CT>
static int run_sepnd_typed(const Plan &P, const void *d_img, int img_dt, void *d_out, cudaStream_t st) {
    int64_t dims[B2F_MAXDIM];
    for (int d = 0; d < B2F_MAXDIM; ++d) dims[d] = P.img_ax.len(d);
    const size_t bytes = (size_t)P.img_ax.count() * sizeof(CT);
    // plan the passes
    struct Pass { const StageInfo *sx, *sy; int yaxis; };
    std::vector<Pass> passes;
    const int na = (int)P.active.size();
    for (int a = 0; a < na;) {
        const StageInfo &s1 = P.stages[P.active[a]];
        if (s1.s->axis == 0 && a + 1 < na && P.stages[P.active[a + 1]].s->axis == 1 &&
            taps_ok_pair(s1.s->len[0], P.stages[P.active[a + 1]].s->len[1])) {
            passes.push_back({&s1, &P.stages[P.active[a + 1]], 1});
            a += 2;
        } else if (s1.s->axis == 0) {
            passes.push_back({&s1, nullptr, 0});
            a += 1;
        } else {
            passes.push_back({nullptr, &s1, s1.s->axis});
            a += 1;
        }
    }
    void *tmp[2] = {nullptr, nullptr};
    const void *src = d_img;
    int src_dt = img_dt;
    CT fill = (CT)P.fill;
    int rc = 0;
    for (size_t i = 0; i < passes.size() && !rc; ++i) {
        const bool last = i + 1 == passes.size();
        void *dst = d_out;
        if (!last) {
            void *&t = tmp[i & 1];
            if (!t) {
                cudaError_t e = cudaMallocAsync(&t, bytes, st);
                if (e != cudaSuccess) { rc = fail(B2F_ENOMEM, "temporary allocation failed: %s", cudaGetErrorString(e)); break; }
            }
            dst = t;
        }
        rc = run_pass<CT>(dims, passes[i].sx, passes[i].sy, passes[i].yaxis, src, src_dt, dst, P.style, fill, 0, 0, 0, 0, st);
        if (passes[i].sx) fill = push_fill<CT>(fill, passes[i].sx->s->taps, passes[i].sx->s->len[0]);
        if (passes[i].sy) fill = push_fill<CT>(fill, passes[i].sy->s->taps, passes[i].sy->s->len[passes[i].yaxis]);
        src = dst;
        src_dt = CtDt<CT>::v;
    }
    for (void *t : tmp)
        if (t) cudaFreeAsync(t, st);
    return rc;
}

// the slab cascade on a gathered buffer of `nplanes` planes whose plane `lo_n` is global plane `slab_first`
template <typename CT>
static int run_slab_typed(const Plan &P, int last, const void *d_src, int img_dt, void *d_out, int64_t nplanes, int64_t lo_n,
                          int64_t own_n, int64_t Zg, int64_t slab_first, cudaStream_t st) {
    int64_t dims[B2F_MAXDIM];
    for (int d = 0; d < B2F_MAXDIM; ++d) dims[d] = P.img_ax.len(d);
    dims[last] = nplanes;
    int64_t plane = 1;
    for (int d = 0; d < last; ++d) plane *= dims[d];
    const size_t bytes = (size_t)(plane * nplanes) * sizeof(CT);
    void *tmp[2] = {nullptr, nullptr};
    const void *src = d_src;
    int src_dt = img_dt;
    CT fill = (CT)P.fill;
    int rc = 0;
    bool sharded_done = false;
    const int na = (int)P.active.size();
    int pass = 0;
    for (int a = 0; a < na && !rc; ++pass) {
        const StageInfo &s1 = P.stages[P.active[a]];
        const StageInfo *sx = nullptr, *sy = nullptr;
        int yaxis = 0, used = 1;
        if (s1.s->axis == 0 && a + 1 < na && P.stages[P.active[a + 1]].s->axis == 1 && last != 1 &&
            taps_ok_pair(s1.s->len[0], P.stages[P.active[a + 1]].s->len[1])) {
            sx = &s1; sy = &P.stages[P.active[a + 1]]; yaxis = 1; used = 2;
        } else if (s1.s->axis == 0) {
            sx = &s1;
        } else {
            sy = &s1; yaxis = s1.s->axis;
        }
        const bool lastpass = a + used == na;
        void *dst = d_out;
        if (!lastpass) {
            void *&t = tmp[pass & 1];
            if (!t) {
                cudaError_t e = cudaMallocAsync(&t, bytes, st);
                if (e != cudaSuccess) { rc = fail(B2F_ENOMEM, "temporary allocation failed: %s", cudaGetErrorString(e)); break; }
            }
            dst = t;
        }
        if (sy && yaxis == last && !sharded_done) {
            rc = run_pass<CT>(dims, sx, sy, yaxis, src, src_dt, dst, P.style, fill, Zg, slab_first - lo_n, lo_n, own_n, st);
            dims[last] = own_n;
            sharded_done = true;
        } else {
            rc = run_pass<CT>(dims, sx, sy, yaxis, src, src_dt, dst, P.style, fill, 0, 0, 0, 0, st);
        }
        if (sx) fill = push_fill<CT>(fill, sx->s->taps, sx->s->len[0]);
        if (sy) fill = push_fill<CT>(fill, sy->s->taps, sy->s->len[yaxis]);
        src = dst;
        src_dt = CtDt<CT>::v;
        a += used;
    }
    for (void *t : tmp)
        if (t) cudaFreeAsync(t, st);
    return rc;
}

int run_sepnd(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    set_path("sepnd");
    return out_dt == B2F_F32 ? run_sepnd_typed<float>(P, d_img, img_dt, d_out, st)
                             : run_sepnd_typed<double>(P, d_img, img_dt, d_out, st);
}

}  // namespace b2f

using namespace b2f;

// Slab form of a separable cascade: the array's LAST axis is sharded.  `img`/`out` hold this rank's owned
// planes [slab_first, slab_first + own_n); halo_lo / halo_hi hold n_halo_lo / n_halo_hi RAW input planes logically
// below / above them (receive buffers, or the neighbour's memory mapped over NVLink).  Semantics = the owned planes of
// b2f_imfilter on the whole array.
static int imfilter_slab_impl(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                              const b2f_border *border, int64_t global_last_dim, int64_t slab_first, const void *halo_lo,
                              int64_t n_halo_lo, const void *halo_hi, int64_t n_halo_hi, const void *flag_lo,
                              const void *flag_hi, int32_t epoch, int32_t lo_early_rows, void *stream,
                              const b2f_slab_xy *xy = nullptr) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !out || !border || !stages) return fail(B2F_EARG, "NULL argument");
    if (xy) {                                   // the halos are present as xy-filtered planes
        if (xy->lo_halo < 0 || xy->lo_own < 0 || xy->hi_own < 0 || xy->hi_halo < 0) return fail(B2F_EARG, "negative plane count");
        if ((xy->lo_halo + xy->lo_own > 0 && !xy->xy_lo) || (xy->hi_own + xy->hi_halo > 0 && !xy->xy_hi))
            return fail(B2F_EARG, "NULL xy buffer");
        n_halo_lo = xy->xy_lo ? xy->lo_halo : 0;
        n_halo_hi = xy->xy_hi ? xy->hi_halo : 0;
        halo_lo = n_halo_lo ? xy->xy_lo : nullptr;        // non-NULL placeholders for the checks below; never read as raw planes
        halo_hi = n_halo_hi ? xy->xy_hi : nullptr;
    }
    if (img->mem != B2F_DEVICE || out->mem != B2F_DEVICE) return fail(B2F_EARG, "b2f_imfilter_slab works on device arrays");
    const int N = img->ndim;
    if (N < 2 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "slab arrays need 2..4 dims and equal rank");
    if (n_halo_lo < 0 || n_halo_hi < 0) return fail(B2F_EARG, "negative halo");
    if ((n_halo_lo > 0 && !halo_lo) || (n_halo_hi > 0 && !halo_hi)) return fail(B2F_EARG, "NULL halo pointer");
    const int last = N - 1;
    const int64_t own_n = img->dims[last];
    for (int d = 0; d < N; ++d)
        if (img->dims[d] != out->dims[d]) return fail(B2F_EDIM, "slab and out extents differ along axis %d", d);
    if (own_n < 1) { set_path("empty"); return 0; }
    if (slab_first < 0 || slab_first + own_n > global_last_dim) return fail(B2F_EDIM, "slab lies outside the global axis");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "slab form supports Float32/Float64 outputs");
    if (border->style > B2F_FILL) return fail(B2F_ENOTSUP, "slab form supports Pad and Fill borders");
    // plan on the GLOBAL array: same validation and border resolution as b2f_imfilter
    b2f_array gi = *img, go = *out;
    gi.dims[last] = go.dims[last] = global_last_dim;
    gi.ptr = go.ptr = nullptr;
    Plan P;
    int rc = make_plan(&gi, &go, stages, nstages, border, nullptr, nullptr, P);
    if (rc) return rc;
    if (P.img_ax.empty()) { set_path("empty"); return 0; }
    if (!sepnd_applicable(P, img->dtype, out->dtype))
        return fail(B2F_ENOTSUP, "slab form takes separable cascades (1-D stages on distinct axes, <= 17 taps)");
    // the planes the owned outputs read along the sharded axis must be present (own or halo), directly or through
    // the global border
    int64_t zlo = 0, zhi = 0;
    for (int a : P.active) { zlo += P.stages[a].lo[last]; zhi += P.stages[a].hi[last]; }
    const int64_t have_lo = slab_first - n_halo_lo, have_hi = slab_first + own_n + n_halo_hi;   // logical [lo, hi)
    for (int64_t z = slab_first + zlo; z < slab_first + own_n + zhi; ++z) {
        if (z >= have_lo && z < have_hi) continue;
        const int64_t g = remap_index(border->style, z, global_last_dim);
        if (g < 0) continue;  // Fill
        if (g < have_lo || g >= have_hi)
            return fail(B2F_EDIM, "halo too small: plane %lld is needed but not present", (long long)g);
    }
    rc = 0;
    if (stream3d_applicable(P, img->dtype, out->dtype)) {
        set_path("stream3d_slab");
        if (xy && (xy->lo_own > own_n || xy->hi_own > own_n)) return fail(B2F_EDIM, "more xy-filtered own planes than the slab has");
        return run_stream3d_slab(P, img->ptr, halo_lo, n_halo_lo, halo_hi, n_halo_hi, slab_first, own_n, out->ptr, st, flag_lo, flag_hi,
                                 epoch, lo_early_rows, xy);
    }
    if (xy) return fail(B2F_ENOTSUP, "xy-filtered halos are available for the fused Float32 3-D kernel only");
    if (flag_lo || flag_hi) {
        return fail(B2F_ENOTSUP, "staged halos are available for the fused Float32 3-D kernel only");
    }
    // general separable cascade: gather [halo_lo; own; halo_hi] once, then one streamed pass per stage; the pass along
    // the sharded axis evaluates the border in GLOBAL plane coordinates and keeps only the owned planes
    set_path("slab");
    const size_t esz = dtype_size(img->dtype);
    int64_t plane = 1;
    for (int d = 0; d < last; ++d) plane *= img->dims[d];
    if (zlo == 0 && zhi == 0) n_halo_lo = n_halo_hi = 0;   // nothing acts along the sharded axis: halos are not read
    int64_t nplanes = n_halo_lo + own_n + n_halo_hi;
    void *ext = nullptr;
    AsyncFrees ext_guard(st);                // released on every exit, error returns included
    const void *src = img->ptr;
    if (n_halo_lo + n_halo_hi > 0) {
        B2F_CUDA(cudaMallocAsync(&ext, (size_t)(plane * nplanes) * esz, st));
        ext_guard.push_back(ext);
        char *e = (char *)ext;
        if (n_halo_lo) B2F_CUDA(cudaMemcpyAsync(e, halo_lo, (size_t)(plane * n_halo_lo) * esz, cudaMemcpyDefault, st));
        B2F_CUDA(cudaMemcpyAsync(e + (size_t)(plane * n_halo_lo) * esz, img->ptr, (size_t)(plane * own_n) * esz, cudaMemcpyDefault, st));
        if (n_halo_hi)
            B2F_CUDA(cudaMemcpyAsync(e + (size_t)(plane * (n_halo_lo + own_n)) * esz, halo_hi, (size_t)(plane * n_halo_hi) * esz, cudaMemcpyDefault, st));
        src = ext;
    }
    rc = out->dtype == B2F_F32
             ? run_slab_typed<float>(P, last, src, img->dtype, out->ptr, nplanes, n_halo_lo, own_n, global_last_dim, slab_first, st)
             : run_slab_typed<double>(P, last, src, img->dtype, out->ptr, nplanes, n_halo_lo, own_n, global_last_dim, slab_first, st);
    return rc;
}

extern "C" {

int b2f_imfilter_slab(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                      const b2f_border *border, int64_t global_last_dim, int64_t slab_first, const void *halo_lo,
                      int64_t n_halo_lo, const void *halo_hi, int64_t n_halo_hi, void *stream) {
    return imfilter_slab_impl(img, out, stages, nstages, border, global_last_dim, slab_first, halo_lo, n_halo_lo, halo_hi,
                              n_halo_hi, nullptr, nullptr, 0, 0, stream);
}

// Staged form: halo_lo / halo_hi are LOCAL buffers that copy engines are still filling when the kernel starts (peer ->
// local copies on another stream, each followed by a one-byte write of `epoch` to flag_lo / flag_hi).  The kernel
// starts at once and only the CTA about to read a halo plane waits for its flag, so the NVLink transfer runs at copy
// speed (no tile over-fetch) underneath the marches instead of stalling them in bursts.  The lower halo, which the first
// wave of CTAs needs at once, comes in two parts: rows [0, lo_early_rows) of every plane first (flag_lo[0]), then the rest
// (flag_lo[1]); a CTA waits for the part its tile rows lie in.
int b2f_imfilter_slab_staged(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                             const b2f_border *border, int64_t global_last_dim, int64_t slab_first, const void *halo_lo,
                             int64_t n_halo_lo, const void *halo_hi, int64_t n_halo_hi, const void *flag_lo,
                             const void *flag_hi, int32_t epoch, int32_t lo_early_rows, void *stream) {
    if ((n_halo_lo > 0 && !flag_lo) || (n_halo_hi > 0 && !flag_hi)) return fail(B2F_EARG, "NULL halo flag");
    if (epoch < 1 || epoch > 255) return fail(B2F_EARG, "epoch must be 1..255");
    return imfilter_slab_impl(img, out, stages, nstages, border, global_last_dim, slab_first, halo_lo, n_halo_lo, halo_hi,
                              n_halo_hi, flag_lo, flag_hi, epoch, lo_early_rows, stream);
}

int b2f_imfilter_slab_xy(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                         const b2f_border *border, int64_t global_last_dim, int64_t slab_first, const b2f_slab_xy *xy,
                         const void *flag_lo, const void *flag_hi, int32_t epoch, int32_t lo_early_rows, void *stream) {
    if (!xy) return fail(B2F_EARG, "NULL argument");
    if ((flag_lo || flag_hi) && (epoch < 1 || epoch > 255)) return fail(B2F_EARG, "epoch must be 1..255");
    return imfilter_slab_impl(img, out, stages, nstages, border, global_last_dim, slab_first, nullptr, 0, nullptr, 0, flag_lo, flag_hi,
                              epoch, lo_early_rows, stream, xy);
}

int b2f_memcpy_async(void *dst, const void *src, uint64_t bytes, void *stream) {
    B2F_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}
int b2f_memcpy2d_async(void *dst, uint64_t dpitch, const void *src, uint64_t spitch, uint64_t width, uint64_t height, void *stream) {
    B2F_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, (cudaStream_t)stream));
    return 0;
}
// stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32, reached through the runtime): a rank signals its
// neighbours by a 32-bit write into their (peer-mapped) flag words and waits on its own — no kernel, no collective
typedef int (*s_memop_fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static s_memop_fn s_memop(const char *name) {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
    return (s_memop_fn)f;
}
int b2f_stream_write32(void *dptr, uint32_t value, void *stream) {
    static s_memop_fn fn = s_memop("cuStreamWriteValue32");
    if (!fn) return fail(B2F_ENOTSUP, "cuStreamWriteValue32 is unavailable");
    if (fn((cudaStream_t)stream, (unsigned long long)(uintptr_t)dptr, value, 0) != 0) return fail(B2F_ECUDA, "cuStreamWriteValue32 failed");
    return 0;
}
int b2f_stream_wait_geq32(void *dptr, uint32_t value, void *stream) {
    static s_memop_fn fn = s_memop("cuStreamWaitValue32");
    if (!fn) return fail(B2F_ENOTSUP, "cuStreamWaitValue32 is unavailable");
    if (fn((cudaStream_t)stream, (unsigned long long)(uintptr_t)dptr, value, 0 /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
        return fail(B2F_ECUDA, "cuStreamWaitValue32 failed");
    return 0;
}
int b2f_memset_async(void *dptr, int32_t byte, uint64_t bytes, void *stream) {
    B2F_CUDA(cudaMemsetAsync(dptr, byte, bytes, (cudaStream_t)stream));
    return 0;
}

}  // extern "C"
