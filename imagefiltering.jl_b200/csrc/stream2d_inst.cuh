// stream2d_inst.cuh — launcher shared by stream2d_f32.cu / stream2d_f64.cu
#pragma once
#include "stream2d.cuh"

namespace b2f {

template <typename IT, typename CT, int LB, int NPL>
static int s2_launch_one(const S2Params<CT, NPL> &P, cudaStream_t st) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int WIN = ((PX + LB - 1 + PX - 1) / PX) * PX;
    constexpr int PW = 32 * PX + WIN;
    const size_t smem = (size_t)S2_WARPS * 2 * 4 * PW * sizeof(CT);
    const long long blocks = (P.nstrips + S2_WARPS - 1) / S2_WARPS;
    if (blocks > 0x7fffffffLL) return fail(B2F_ENOTSUP, "stream2d grid too large");
    stream2d_kernel<IT, CT, LB, NPL><<<(unsigned)blocks, S2_WARPS * 32, smem, st>>>(P);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

template <typename IT, typename CT, int NPL>
static int s2_launch_lb(const S2Params<CT, NPL> &P, cudaStream_t st) {
    const int L = P.Lx > P.Ly ? P.Lx : P.Ly;
    if (L <= 4) return s2_launch_one<IT, CT, 4, NPL>(P, st);
    if (L <= 8) return s2_launch_one<IT, CT, 8, NPL>(P, st);
    if constexpr (NPL == 1) return s2_launch_one<IT, CT, 16, 1>(P, st);
    else return fail(B2F_ENOTSUP, "stream2d: two planes need <= 8 taps");
}

template <typename CT, int NPL>
static int s2_launch_it(const S2Params<CT, NPL> &P, int img_dt, cudaStream_t st) {
    switch (img_dt) {
        case B2F_U8: case B2F_N0F8: return s2_launch_lb<uint8_t, CT, NPL>(P, st);
        case B2F_F32: return s2_launch_lb<float, CT, NPL>(P, st);
        case B2F_F64: return s2_launch_lb<double, CT, NPL>(P, st);
    }
    return fail(B2F_ENOTSUP, "stream2d: unsupported image dtype");
}

template <typename CT, int NPL> int launch_stream2d(const S2Params<CT, NPL> &P, int img_dt, cudaStream_t st);

}  // namespace b2f
