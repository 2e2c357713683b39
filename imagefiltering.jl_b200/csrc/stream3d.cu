// stream3d.cu — applicability, parameter set-up and launch of the fused 3-D separable kernel (stream3d.cuh)
#include <cstdlib>

#include "stream3d.cuh"

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

template <int LXT, int LYT, int LZT>
static int s3_launch_one(const S3Params &P, long long nblocks, cudaStream_t st) {
    constexpr int LBY = LYT ? LYT : S3_MAXTAPS;
    constexpr int RH = S3_T + LBY - 1;
    const size_t smem = sizeof(float) * (size_t)(S3_NRAW * RH * S3_RWP + S3_NXF * RH * S3_XFP + S3_RING * S3_T * S3_T) + sizeof(int) * RH;
    auto kern = stream3d_kernel<LXT, LYT, LZT>;
    static thread_local bool configured = false;
    if (!configured) {
        B2F_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    kern<<<(unsigned)nblocks, S3_NT, smem, st>>>(P);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

// `own` holds planes [own_first, own_first+own_n) of a volume with Zg planes; lo/hi hold lo_n/hi_n planes below/above.
int run_stream3d_slab(const Plan &Pl, const void *own, const void *lo, int64_t lo_n, const void *hi, int64_t hi_n,
                      int64_t own_first, int64_t own_n, void *d_out, cudaStream_t st) {
    S3Params P;
    memset(&P, 0, sizeof P);
    P.own = (const float *)own; P.lo = (const float *)lo; P.hi = (const float *)hi;
    P.own_first = (int)own_first; P.own_n = (int)own_n; P.lo_n = (int)lo_n; P.hi_n = (int)hi_n;
    P.Zg = (int)Pl.img_ax.len(2);
    P.W = (int)Pl.img_ax.len(0); P.H = (int)Pl.img_ax.len(1);
    P.plane = (long long)P.W * P.H;
    P.out = (float *)d_out;
    P.style = Pl.style; P.fill = (float)Pl.fill;
    const StageInfo &sx = Pl.stages[Pl.active[0]], &sy = Pl.stages[Pl.active[1]], &sz = Pl.stages[Pl.active[2]];
    P.Lx = (int)sx.s->len[0]; P.klox = (int)sx.lo[0];
    P.Ly = (int)sy.s->len[1]; P.kloy = (int)sy.lo[1];
    P.Lz = (int)sz.s->len[2]; P.kloz = (int)sz.lo[2];
    for (int j = 0; j < P.Lx; ++j) P.kx[j] = (float)sx.s->taps[j];
    for (int j = 0; j < P.Ly; ++j) P.ky[j] = (float)sy.s->taps[j];
    for (int j = 0; j < P.Lz; ++j) P.kz[j] = (float)sz.s->taps[j];
    auto al16 = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    P.vec_in = (P.W % 4 == 0) && al16(own) && (lo_n == 0 || al16(lo)) && (hi_n == 0 || al16(hi));
    P.vec_out = (P.W % 2 == 0) && reinterpret_cast<uintptr_t>(d_out) % 8 == 0;
    P.ntx = (P.W + S3_T - 1) / S3_T;
    P.nty = (P.H + S3_T - 1) / S3_T;
    // z-chunks: one march per tile unless the xy tiling alone cannot fill the machine (each chunk re-runs Lz-1 planes)
    const long long tiles = (long long)P.ntx * P.nty;
    long long nch = 1;
    if (tiles < 2 * 148) {
        nch = (2 * 148 + tiles - 1) / tiles;
        const long long maxch = own_n / 32 > 0 ? own_n / 32 : 1;
        if (nch > maxch) nch = maxch;
    }
    P.zchunk = (int)((own_n + nch - 1) / nch);
    nch = (own_n + P.zchunk - 1) / P.zchunk;
    const long long nblocks = tiles * nch;
    if (nblocks > 0x7fffffffLL) return fail(B2F_ENOTSUP, "stream3d grid too large");
    if (P.Lx == 17 && P.Ly == 17 && P.Lz == 17) return s3_launch_one<17, 17, 17>(P, nblocks, st);
    if (P.Lx == 9 && P.Ly == 9 && P.Lz == 9) return s3_launch_one<9, 9, 9>(P, nblocks, st);
    if (P.Lx == 5 && P.Ly == 5 && P.Lz == 5) return s3_launch_one<5, 5, 5>(P, nblocks, st);
    if (P.Lx == 3 && P.Ly == 3 && P.Lz == 3) return s3_launch_one<3, 3, 3>(P, nblocks, st);
    return s3_launch_one<0, 0, 0>(P, nblocks, st);
}

int run_stream3d(const Plan &P, const void *d_img, void *d_out, cudaStream_t st) {
    set_path("stream3d");
    return run_stream3d_slab(P, d_img, nullptr, 0, nullptr, 0, 0, P.img_ax.len(2), d_out, st);
}

}  // namespace b2f
