// fused2d.cu — applicability test and parameter set-up of the fused 2-D separable path (fused2d.cuh)
#include "fused2d.cuh"

namespace b2f {

template <> int launch_fused2d<float, 1>(F2Params<float, 1> &, bool, int, cudaStream_t);
template <> int launch_fused2d<float, 2>(F2Params<float, 2> &, bool, int, cudaStream_t);
template <> int launch_fused2d<double, 1>(F2Params<double, 1> &, bool, int, cudaStream_t);
template <> int launch_fused2d<double, 2>(F2Params<double, 2> &, bool, int, cudaStream_t);

static bool same_box(const Box &a, const Box &b) {
    for (int d = 0; d < B2F_MAXDIM; ++d)
        if (a.lo[d] != b.lo[d] || a.hi[d] != b.hi[d]) return false;
    return true;
}

// Two active 1-D stages, one on axis 0 and one on axis 1 (any order, any offsets), same shape for
// every plane; float or double output; batch axes untouched.
bool fused2d_applicable(const Plan *plans, int nplanes, int img_dt, const int *out_dt) {
    if (nplanes < 1 || nplanes > 2) return false;
    const Plan &P0 = plans[0];
    if (P0.ndim < 2) return false;
    if (out_dt[0] != B2F_F32 && out_dt[0] != B2F_F64) return false;
    if (P0.style > B2F_NOPAD) return false;
    if (P0.img_ax.len(0) >= (1LL << 30) || P0.img_ax.len(1) >= (1LL << 30)) return false;
    for (int p = 0; p < nplanes; ++p) {
        const Plan &P = plans[p];
        if (out_dt[p] != out_dt[0]) return false;
        if (P.active.size() != 2) return false;
        if (!same_box(P.out_ax, P0.out_ax) || !same_box(P.roi, P0.roi)) return false;
        int axes[2];
        for (int a = 0; a < 2; ++a) {
            const StageInfo &si = P.stages[P.active[a]];
            if (si.s->kind != B2F_STAGE_1D) return false;
            axes[a] = si.s->axis;
            if (si.s->len[si.s->axis] > F2_MAXTAPS) return false;
            const StageInfo &s0 = P0.stages[P0.active[a]];
            if (s0.s->axis != si.s->axis || s0.lo[si.s->axis] != si.lo[si.s->axis] || s0.hi[si.s->axis] != si.hi[si.s->axis])
                return false;
        }
        if (!((axes[0] == 0 && axes[1] == 1) || (axes[0] == 1 && axes[1] == 0))) return false;
        for (int d = 2; d < B2F_MAXDIM; ++d)   // batch axes: identical, fully covered
            if (P.roi.lo[d] != P.img_ax.lo[d] || P.roi.hi[d] != P.img_ax.hi[d] || P.out_ax.lo[d] != P.img_ax.lo[d] ||
                P.out_ax.hi[d] != P.img_ax.hi[d])
                return false;
    }
    (void)img_dt;
    return true;
}

template <typename CT, int NPL>
static int run_typed(const Plan *plans, const void *d_img, int img_dt, void *const *d_outs, cudaStream_t st) {
    const Plan &P0 = plans[0];
    F2Params<CT, NPL> P;
    memset(&P, 0, sizeof P);
    P.img = d_img; P.img_dt = img_dt;
    P.W = (int)P0.img_ax.len(0); P.H = (int)P0.img_ax.len(1);
    P.img_plane = (long long)P.W * P.H;
    P.out_pitch = P0.out_ax.len(0);
    P.out_plane = P0.out_ax.len(0) * P0.out_ax.len(1);
    P.out_ox = (int)(P0.out_ax.lo[0] - P0.img_ax.lo[0]);
    P.out_oy = (int)(P0.out_ax.lo[1] - P0.img_ax.lo[1]);
    P.rx0 = (int)(P0.roi.lo[0] - P0.img_ax.lo[0]); P.ry0 = (int)(P0.roi.lo[1] - P0.img_ax.lo[1]);
    P.rw = (int)P0.roi.len(0); P.rh = (int)P0.roi.len(1);
    P.style = P0.style; P.fill = (CT)P0.fill;
    const bool xfirst = P0.stages[P0.active[0]].s->axis == 0;
    for (int p = 0; p < NPL; ++p) {
        P.out[p] = d_outs[p];
        for (int a = 0; a < 2; ++a) {
            const StageInfo &si = plans[p].stages[plans[p].active[a]];
            const int ax = si.s->axis;
            const int L = (int)si.s->len[ax];
            CT *dst = ax == 0 ? P.kx[p] : P.ky[p];
            for (int j = 0; j < L; ++j) dst[j] = (CT)si.s->taps[j];
            if (ax == 0) { P.Lx = L; P.klox = (int)si.lo[0]; } else { P.Ly = L; P.kloy = (int)si.lo[1]; }
        }
    }
    const long long nbatch = P0.img_ax.len(2) * P0.img_ax.len(3);
    return launch_fused2d<CT, NPL>(P, xfirst, (int)nbatch, st);
}

int run_fused2d(const Plan *plans, int nplanes, const void *d_img, int img_dt, void *const *d_outs,
                const int *out_dt, cudaStream_t st) {
    set_path(nplanes == 1 ? "fused2d" : "fused2d_grad");
    if (out_dt[0] == B2F_F32)
        return nplanes == 1 ? run_typed<float, 1>(plans, d_img, img_dt, d_outs, st)
                            : run_typed<float, 2>(plans, d_img, img_dt, d_outs, st);
    return nplanes == 1 ? run_typed<double, 1>(plans, d_img, img_dt, d_outs, st)
                        : run_typed<double, 2>(plans, d_img, img_dt, d_outs, st);
}

}  // namespace b2f
