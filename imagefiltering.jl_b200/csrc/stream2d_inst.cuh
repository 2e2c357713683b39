// stream2d_inst.cuh — launcher shared by the stream2d_<in>_<ct>.cu instantiation files
#pragma once
#include <type_traits>

#include "stream2d.cuh"

namespace b2f {

template <typename IT, typename CT, int LXT, int LYT, int LB, int NPL, int RB, int ROT, bool XS = true, bool YS = true, bool FMA = false>
static int s2_launch_one(const S2Params<CT, NPL> &P, cudaStream_t st) {
    constexpr int PX = S2Vec<CT>::PX;
    constexpr int LBX = XS ? (LXT ? LXT : LB) : 1;
    constexpr int WIN = ((PX + LBX - 1 + PX - 1) / PX) * PX;
    constexpr int PW = 32 * PX + WIN;
    constexpr int LBY = YS ? (LYT ? LYT : LB) : 1;
    const size_t smem = (size_t)S2_WARPS * 2 * RB * PW * sizeof(CT);
    S2Params<CT, NPL> Q = P;                       // right-align the y taps in this instantiation's LBY slots
    if (YS) {
        if (P.Ly > LBY) return fail(B2F_ENOTSUP, "stream2d: more y taps than the instantiation holds");
        for (int p = 0; p < NPL; ++p) {
            for (int j = 0; j < S2_MAXTAPS; ++j) Q.kyt[p][j] = (CT)0;
            for (int j = 0; j < P.Ly; ++j) Q.kyt[p][LBY - P.Ly + j] = P.ky[p][j];
        }
    }
    const long long blocks = (P.nstrips + S2_WARPS - 1) / S2_WARPS;
    if (blocks > 0x7fffffffLL) return fail(B2F_ENOTSUP, "stream2d grid too large");
    stream2d_kernel<IT, CT, LXT, LYT, LB, NPL, RB, ROT, XS, YS, FMA><<<(unsigned)blocks, S2_WARPS * 32, smem, st>>>(Q);
    count_launch();
    B2F_CUDA(cudaGetLastError());
    return 0;
}

// hot tap counts get exact instantiations; everything else runs in a bucket with run-time counts
template <typename IT, typename CT, int NPL>
static int s2_launch(const S2Params<CT, NPL> &P, cudaStream_t st) {
    const int Lx = P.Lx, Ly = P.Ly;
    const int L = Lx > Ly ? Lx : Ly;
    // accum mode B2F_ACCUM_FMA is a permission, not an obligation: the fused form exists for Float64 compute with the hot tap
    // counts (the reference-typed configs C1 = 13 x 13 and C2 = two 3 x 3 planes, and their neighbours)
    if constexpr (sizeof(CT) == 8) {
        if (P.fma) {
            if (Lx == 3 && Ly == 3) return s2_launch_one<IT, CT, 3, 3, 4, NPL, 6, 3, true, true, true>(P, st);
            if constexpr (NPL == 1) {
                if (Lx == 5 && Ly == 5) return s2_launch_one<IT, CT, 5, 5, 8, 1, 3, 6, true, true, true>(P, st);
                if (Lx == 7 && Ly == 7) return s2_launch_one<IT, CT, 7, 7, 8, 1, 4, 8, true, true, true>(P, st);
                if (Lx == 9 && Ly == 9) return s2_launch_one<IT, CT, 9, 9, 16, 1, 3, 9, true, true, true>(P, st);
                if (Lx == 13 && Ly == 13) return s2_launch_one<IT, CT, 13, 13, 16, 1, 4, 16, true, true, true>(P, st);
                if (Lx == 17 && Ly == 17) return s2_launch_one<IT, CT, 17, 17, 20, 1, 3, 18, true, true, true>(P, st);
            }
        }
    }
    if (Lx == 3 && Ly == 3) return s2_launch_one<IT, CT, 3, 3, 4, NPL, 6, 3>(P, st);
    if constexpr (NPL == 1) {
        if (Lx == 5 && Ly == 5) return s2_launch_one<IT, CT, 5, 5, 8, 1, 3, 6>(P, st);
        if (Lx == 7 && Ly == 7) return s2_launch_one<IT, CT, 7, 7, 8, 1, 4, 8>(P, st);
        if (Lx == 9 && Ly == 9) return s2_launch_one<IT, CT, 9, 9, 16, 1, 3, 9>(P, st);
        if (Lx == 13 && Ly == 13) return s2_launch_one<IT, CT, 13, 13, 16, 1, 4, 16>(P, st);
        if (Lx == 17 && Ly == 17) return s2_launch_one<IT, CT, 17, 17, 20, 1, 3, 18>(P, st);
    }
    if (L <= 4) return s2_launch_one<IT, CT, 0, 0, 4, NPL, 4, 4>(P, st);
    if (L <= 8) return s2_launch_one<IT, CT, 0, 0, 8, NPL, 4, 8>(P, st);
    if constexpr (NPL == 1) {
        if (L <= 16) return s2_launch_one<IT, CT, 0, 0, 16, 1, 4, 16>(P, st);
    }
    return fail(B2F_ENOTSUP, "stream2d: tap count outside the instantiated range");
}

// a single 1-D stage: along x only (no y stage) or along y only (no x stage), one plane
template <typename IT, typename CT>
static int s2_launch_single(const S2Params<CT, 1> &P, bool along_x, cudaStream_t st) {
    if (along_x) {
        const int L = P.Lx;
        if constexpr (std::is_same<IT, CT>::value) {
            if (L == 13) return s2_launch_one<IT, CT, 13, 0, 16, 1, 8, 1, true, false>(P, st);
            if (L == 17) return s2_launch_one<IT, CT, 17, 0, 20, 1, 8, 1, true, false>(P, st);
        }
        if (L <= 4) return s2_launch_one<IT, CT, 0, 0, 4, 1, 8, 1, true, false>(P, st);
        if (L <= 8) return s2_launch_one<IT, CT, 0, 0, 8, 1, 8, 1, true, false>(P, st);
        if (L <= 16) return s2_launch_one<IT, CT, 0, 0, 16, 1, 8, 1, true, false>(P, st);
    } else {
        const int L = P.Ly;
        if constexpr (std::is_same<IT, CT>::value) {
            if (L == 13) return s2_launch_one<IT, CT, 0, 13, 16, 1, 4, 16, false, true>(P, st);
            if (L == 17) return s2_launch_one<IT, CT, 0, 17, 20, 1, 3, 18, false, true>(P, st);
        }
        if (L <= 4) return s2_launch_one<IT, CT, 0, 0, 4, 1, 4, 4, false, true>(P, st);
        if (L <= 8) return s2_launch_one<IT, CT, 0, 0, 8, 1, 4, 8, false, true>(P, st);
        if (L <= 16) return s2_launch_one<IT, CT, 0, 0, 16, 1, 4, 16, false, true>(P, st);
    }
    return fail(B2F_ENOTSUP, "stream2d: tap count outside the instantiated range");
}

// one definition per (input type, compute type) lives in its own .cu file so they compile in parallel
template <typename IT, typename CT, int NPL> int launch_stream2d(const S2Params<CT, NPL> &P, cudaStream_t st);
template <typename IT, typename CT> int launch_stream1d(const S2Params<CT, 1> &P, bool along_x, cudaStream_t st);

#define B2F_S2_INSTANTIATE(IT, CT)                                                                                  \
    template <> int launch_stream2d<IT, CT, 1>(const S2Params<CT, 1> &P, cudaStream_t st) { return s2_launch<IT, CT, 1>(P, st); } \
    template <> int launch_stream2d<IT, CT, 2>(const S2Params<CT, 2> &P, cudaStream_t st) { return s2_launch<IT, CT, 2>(P, st); } \
    template <> int launch_stream1d<IT, CT>(const S2Params<CT, 1> &P, bool ax, cudaStream_t st) { return s2_launch_single<IT, CT>(P, ax, st); }

}  // namespace b2f
