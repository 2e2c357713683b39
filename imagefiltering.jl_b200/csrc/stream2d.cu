// stream2d.cu — applicability test and parameter set-up of the warp-streamed 2-D separable path
#include "stream2d.cuh"

namespace b2f {

template <typename IT, typename CT, int NPL> int launch_stream2d(const S2Params<CT, NPL> &P, cudaStream_t st);
#define B2F_S2_DECL(IT, CT)                                                                  \
    template <> int launch_stream2d<IT, CT, 1>(const S2Params<CT, 1> &, cudaStream_t);       \
    template <> int launch_stream2d<IT, CT, 2>(const S2Params<CT, 2> &, cudaStream_t);
B2F_S2_DECL(uint8_t, float)
B2F_S2_DECL(uint8_t, double)
B2F_S2_DECL(float, float)
B2F_S2_DECL(float, double)
B2F_S2_DECL(double, float)
B2F_S2_DECL(double, double)

// On top of fused2d's conditions: stage order x then y, u8/N0f8/f32/f64 input, <= 16 taps (<= 8 for two planes).
bool stream2d_applicable(const Plan *plans, int nplanes, int img_dt, const int *out_dt) {
    if (!fused2d_applicable(plans, nplanes, img_dt, out_dt)) return false;
    if (img_dt != B2F_U8 && img_dt != B2F_N0F8 && img_dt != B2F_F32 && img_dt != B2F_F64) return false;
    const Plan &P0 = plans[0];
    const StageInfo &s1 = P0.stages[P0.active[0]], &s2 = P0.stages[P0.active[1]];
    if (s1.s->axis != 0 || s2.s->axis != 1) return false;
    const int64_t Lx = s1.s->len[0], Ly = s2.s->len[1];
    const int64_t L = Lx > Ly ? Lx : Ly;
    if (nplanes == 1 && Lx == 17 && Ly == 17) return true;
    if (L > (nplanes == 1 ? 16 : 8)) return false;
    return true;
}

template <typename CT, int NPL>
static int run_typed(const Plan *plans, const void *d_img, int img_dt, void *const *d_outs, cudaStream_t st) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int CW = 32 * PX;
    const Plan &P0 = plans[0];
    S2Params<CT, NPL> P;
    memset(&P, 0, sizeof P);
    P.img = d_img;
    P.n0_r = img_dt == B2F_N0F8 ? (CT)1 / (CT)255 : (CT)1;
    P.n0_c = img_dt == B2F_N0F8 ? (CT)255 : (CT)1;
    P.W = (int)P0.img_ax.len(0); P.H = (int)P0.img_ax.len(1);
    P.Hg = P.H; P.y_first = 0;
    P.img_plane = (long long)P.W * P.H;
    P.out_pitch = P0.out_ax.len(0);
    P.out_plane = P0.out_ax.len(0) * P0.out_ax.len(1);
    P.out_ox = (int)(P0.out_ax.lo[0] - P0.img_ax.lo[0]);
    P.out_oy = (int)(P0.out_ax.lo[1] - P0.img_ax.lo[1]);
    P.rx0 = (int)(P0.roi.lo[0] - P0.img_ax.lo[0]); P.ry0 = (int)(P0.roi.lo[1] - P0.img_ax.lo[1]);
    P.rw = (int)P0.roi.len(0); P.rh = (int)P0.roi.len(1);
    P.style = P0.style; P.fill = (CT)P0.fill;
    P.fma = accum_mode() == B2F_ACCUM_FMA;
    bool aligned = (P.out_pitch % PX == 0) && (P.out_plane % PX == 0) && ((P.rx0 - P.out_ox) % PX == 0);
    for (int p = 0; p < NPL; ++p) {
        P.out[p] = d_outs[p];
        aligned = aligned && (reinterpret_cast<uintptr_t>(d_outs[p]) % 16 == 0);
        const StageInfo &sx = plans[p].stages[plans[p].active[0]], &sy = plans[p].stages[plans[p].active[1]];
        P.Lx = (int)sx.s->len[0]; P.klox = (int)sx.lo[0];
        P.Ly = (int)sy.s->len[1]; P.kloy = (int)sy.lo[1];
        for (int j = 0; j < P.Lx; ++j) P.kx[p][j] = (CT)sx.s->taps[j];
        for (int j = 1; j < P.Lx; ++j) P.kxp[p][j] = make_float2((float)sx.s->taps[j], (float)sx.s->taps[j - 1]);
        for (int j = 0; j < P.Ly; ++j) P.ky[p][j] = (CT)sy.s->taps[j];
    }
    P.vec_ok = aligned ? 1 : 0;
    const long long nbatch = P0.img_ax.len(2) * P0.img_ax.len(3);
    P.nsx = (P.rw + CW - 1) / CW;
    // strip height: as tall as possible (less y-halo re-read) while still filling the machine with warps
    const long long want = (long long)sm_count() * 16 * 6;
    int SH = 256;
    while (SH > 32 && (long long)P.nsx * ((P.rh + SH - 1) / SH) * nbatch < want) SH >>= 1;
    P.SH = SH;
    P.nsy = (P.rh + SH - 1) / SH;
    P.nstrips = (long long)P.nsx * P.nsy * nbatch;
    switch (img_dt) {
        case B2F_U8: case B2F_N0F8: return launch_stream2d<uint8_t, CT, NPL>(P, st);
        case B2F_F32: return launch_stream2d<float, CT, NPL>(P, st);
        default: return launch_stream2d<double, CT, NPL>(P, st);
    }
}

int run_stream2d(const Plan *plans, int nplanes, const void *d_img, int img_dt, void *const *d_outs,
                 const int *out_dt, cudaStream_t st) {
    set_path(nplanes == 1 ? "stream2d" : "stream2d_grad");
    if (out_dt[0] == B2F_F32)
        return nplanes == 1 ? run_typed<float, 1>(plans, d_img, img_dt, d_outs, st)
                            : run_typed<float, 2>(plans, d_img, img_dt, d_outs, st);
    return nplanes == 1 ? run_typed<double, 1>(plans, d_img, img_dt, d_outs, st)
                        : run_typed<double, 2>(plans, d_img, img_dt, d_outs, st);
}

}  // namespace b2f
