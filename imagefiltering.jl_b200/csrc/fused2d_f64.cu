// fused2d_f64.cu — double (bit-exact mode) instantiations of the fused 2-D separable kernel (see fused2d.cuh)
#include "fused2d.cuh"
namespace b2f {
template <> int launch_fused2d<double, 1>(F2Params<double, 1> &P, bool xfirst, int nbatch, cudaStream_t st) {
    return launch_fused2d_impl<double, 1>(P, xfirst, nbatch, st);
}
template <> int launch_fused2d<double, 2>(F2Params<double, 2> &P, bool xfirst, int nbatch, cudaStream_t st) {
    return launch_fused2d_impl<double, 2>(P, xfirst, nbatch, st);
}
}  // namespace b2f
