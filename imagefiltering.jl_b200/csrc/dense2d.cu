// dense2d.cu — applicability test and parameter set-up of the dense 2-D path (dense2d.cuh)
#include "dense2d.cuh"

namespace b2f {

// exactly one active stage, dense over axes 0/1 (later axes are a batch), <= 32 x 64 taps, float/double output
bool dense2d_applicable(const Plan &P, int img_dt, int out_dt) {
    if (P.ndim < 2 || P.active.size() != 1) return false;
    if (out_dt != B2F_F32 && out_dt != B2F_F64) return false;
    const StageInfo &si = P.stages[P.active[0]];
    if (si.s->kind != B2F_STAGE_DENSE) return false;
    for (int d = 2; d < B2F_MAXDIM; ++d) {
        if (si.lo[d] != 0 || si.hi[d] != 0) return false;
        if (P.roi.lo[d] != P.img_ax.lo[d] || P.roi.hi[d] != P.img_ax.hi[d] || P.out_ax.lo[d] != P.img_ax.lo[d] ||
            P.out_ax.hi[d] != P.img_ax.hi[d])
            return false;
    }
    const int64_t Kx = si.hi[0] - si.lo[0] + 1, Ky = si.hi[1] - si.lo[1] + 1;
    if (Kx > 32 || Ky > D2_MAXKY || Kx * Ky < 2) return false;
    if (P.img_ax.len(0) >= (1LL << 30) || P.img_ax.len(1) >= (1LL << 30)) return false;
    (void)img_dt;
    return true;
}

template <typename CT>
static int run_typed(const Plan &P0, const void *d_img, int img_dt, void *d_out, cudaStream_t st,
                     int (*launch)(D2Params<CT> &, int, cudaStream_t)) {
    constexpr int R = D2Vec<CT>::N;
    const StageInfo &si = P0.stages[P0.active[0]];
    D2Params<CT> P;
    memset(&P, 0, sizeof P);
    P.img = d_img; P.img_dt = img_dt;
    P.W = (int)P0.img_ax.len(0); P.H = (int)P0.img_ax.len(1);
    P.img_plane = (long long)P.W * P.H;
    P.out = d_out;
    P.out_pitch = P0.out_ax.len(0);
    P.out_plane = P0.out_ax.len(0) * P0.out_ax.len(1);
    P.out_ox = (int)(P0.out_ax.lo[0] - P0.img_ax.lo[0]);
    P.out_oy = (int)(P0.out_ax.lo[1] - P0.img_ax.lo[1]);
    P.rx0 = (int)(P0.roi.lo[0] - P0.img_ax.lo[0]); P.ry0 = (int)(P0.roi.lo[1] - P0.img_ax.lo[1]);
    P.rw = (int)P0.roi.len(0); P.rh = (int)P0.roi.len(1);
    P.style = P0.style; P.fill = (CT)P0.fill;
    P.Kx = (int)(si.hi[0] - si.lo[0] + 1); P.Ky = (int)(si.hi[1] - si.lo[1] + 1);
    P.klox = (int)si.lo[0]; P.kloy = (int)si.lo[1];
    const int KXP = ((P.Kx + R - 1) / R) * R;
    std::vector<CT> h((size_t)P.Ky * KXP, (CT)0);
    for (int J = 0; J < P.Ky; ++J)
        for (int j = 0; j < P.Kx; ++j) h[(size_t)J * KXP + j] = (CT)si.s->taps[(size_t)J * P.Kx + j];
    CT *d_taps = nullptr;
    B2F_CUDA(cudaMallocAsync((void **)&d_taps, h.size() * sizeof(CT), st));
    cudaError_t e = cudaMemcpyAsync(d_taps, h.data(), h.size() * sizeof(CT), cudaMemcpyHostToDevice, st);
    int rc = 0;
    if (e != cudaSuccess) rc = fail(B2F_ECUDA, "tap upload failed: %s", cudaGetErrorString(e));
    P.taps = d_taps;
    const long long nbatch = P0.img_ax.len(2) * P0.img_ax.len(3);
    if (!rc) rc = launch(P, (int)nbatch, st);
    cudaFreeAsync(d_taps, st);
    return rc;
}

int run_dense2d(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    set_path("dense2d");
    if (out_dt == B2F_F32) return run_typed<float>(P, d_img, img_dt, d_out, st, launch_dense2d_f32);
    return run_typed<double>(P, d_img, img_dt, d_out, st, launch_dense2d_f64);
}

}  // namespace b2f
