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

static bool taps_ok_single(int64_t L) { return L >= 1 && (L <= 16 || L == 17); }
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
    const long long want = 148LL * 16 * 6;
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
    P.Lx = 1; P.Ly = 1;
    if (sx) {
        P.Lx = (int)sx->s->len[0]; P.klox = (int)sx->lo[0];
        for (int j = 0; j < P.Lx; ++j) P.kx[0][j] = (CT)sx->s->taps[j];
    }
    if (sy) {
        P.Ly = (int)sy->s->len[yaxis]; P.kloy = (int)sy->lo[yaxis];
        for (int d = 0; d < P.Ly; ++d) P.kyr[0][d] = (CT)sy->s->taps[P.Ly - 1 - d];
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

int run_sepnd(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    set_path("sepnd");
    return out_dt == B2F_F32 ? run_sepnd_typed<float>(P, d_img, img_dt, d_out, st)
                             : run_sepnd_typed<double>(P, d_img, img_dt, d_out, st);
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_imfilter_slab(const b2f_array *img, const b2f_array *out, const b2f_stage *stages, int32_t nstages,
                                 const b2f_border *border, int64_t global_last_dim, int64_t slab_first,
                                 int64_t halo_lo, int64_t halo_hi, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !out || !border || !stages) return fail(B2F_EARG, "NULL argument");
    if (img->mem != B2F_DEVICE || out->mem != B2F_DEVICE) return fail(B2F_EARG, "b2f_imfilter_slab works on device arrays");
    const int N = img->ndim;
    if (N < 2 || N > B2F_MAXDIM || out->ndim != N) return fail(B2F_EDIM, "slab arrays need 2..4 dims and equal rank");
    if (halo_lo < 0 || halo_hi < 0) return fail(B2F_EARG, "negative halo");
    const int last = N - 1;
    const int64_t owned = img->dims[last] - halo_lo - halo_hi;
    if (owned < 1 || out->dims[last] != owned) return fail(B2F_EDIM, "out must hold exactly the owned planes");
    for (int d = 0; d < last; ++d)
        if (img->dims[d] != out->dims[d]) return fail(B2F_EDIM, "slab and out extents differ along axis %d", d);
    if (slab_first < 0 || slab_first + owned > global_last_dim) return fail(B2F_EDIM, "slab lies outside the global axis");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_ENOTSUP, "slab form supports Float32/Float64 outputs");
    if (img->dtype != out->dtype) return fail(B2F_ENOTSUP, "slab form expects the intermediate in eltype(out)");
    if (border->style > B2F_FILL) return fail(B2F_ENOTSUP, "slab form supports Pad and Fill borders");
    // exactly one non-copy stage, 1-D along the sharded (last) axis
    const b2f_stage *zs = nullptr;
    for (int s = 0; s < nstages; ++s) {
        const b2f_stage &t = stages[s];
        if (t.kind != B2F_STAGE_1D) return fail(B2F_ENOTSUP, "slab form takes 1-D stages");
        if (t.axis < 0 || t.axis >= N) return fail(B2F_EDIM, "stage axis out of range");
        const bool copy = t.len[t.axis] == 1 && t.lo[t.axis] == 0 && t.taps && t.taps[0] == 1.0;
        if (copy) continue;
        if (t.axis != last || zs) return fail(B2F_ENOTSUP, "slab form runs the single stage along the sharded axis; run the other stages with b2f_imfilter first");
        zs = &t;
    }
    if (!zs) return fail(B2F_EARG, "no stage along the sharded axis");
    const int64_t L = zs->len[last], klo = zs->lo[last];
    if (!(L <= 16 || L == 17)) return fail(B2F_ENOTSUP, "slab form supports up to 16 (or 17) taps");
    const int64_t H = img->dims[last], y_first = slab_first - halo_lo;
    // every plane the owned outputs read must be in the buffer, directly or through the global border
    for (int64_t i = halo_lo + klo; i <= halo_lo + owned - 1 + klo + L - 1; ++i) {
        if (i >= 0 && i < H) continue;
        const int64_t g = remap_index(border->style, i + y_first, global_last_dim);
        if (g < 0) continue;  // Fill
        if (g - y_first < 0 || g - y_first >= H) return fail(B2F_EDIM, "halo too small: plane %lld is needed but not present", (long long)g);
    }
    StageInfo si;
    si.s = zs;
    for (int d = 0; d < B2F_MAXDIM; ++d) si.lo[d] = si.hi[d] = 0;
    si.lo[last] = klo; si.hi[last] = klo + L - 1; si.copy = false;
    int64_t dims[B2F_MAXDIM];
    for (int d = 0; d < B2F_MAXDIM; ++d) dims[d] = d < N ? img->dims[d] : 1;
    set_path("slab");
    if (out->dtype == B2F_F32)
        return run_pass<float>(dims, nullptr, &si, last, img->ptr, B2F_F32, out->ptr, border->style, (float)border->fill,
                               global_last_dim, y_first, halo_lo, owned, st);
    return run_pass<double>(dims, nullptr, &si, last, img->ptr, B2F_F64, out->ptr, border->style, border->fill,
                            global_last_dim, y_first, halo_lo, owned, st);
}
