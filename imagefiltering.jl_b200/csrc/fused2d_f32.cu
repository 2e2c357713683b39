// fused2d_f32.cu — float instantiations of the fused 2-D separable kernel (see fused2d.cuh)
#include "fused2d.cuh"
namespace b2f {
template <> int launch_fused2d<float, 1>(F2Params<float, 1> &P, bool xfirst, int nbatch, cudaStream_t st) {
    return launch_fused2d_impl<float, 1>(P, xfirst, nbatch, st);
}
template <> int launch_fused2d<float, 2>(F2Params<float, 2> &P, bool xfirst, int nbatch, cudaStream_t st) {
    return launch_fused2d_impl<float, 2>(P, xfirst, nbatch, st);
}
}  // namespace b2f
