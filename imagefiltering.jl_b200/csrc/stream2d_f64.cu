// stream2d_f64.cu — double (bit-exact mode) instantiations of the warp-streamed fused 2-D separable kernel
#include "stream2d_inst.cuh"
namespace b2f {
template <> int launch_stream2d<double, 1>(const S2Params<double, 1> &P, int dt, cudaStream_t st) { return s2_launch_it<double, 1>(P, dt, st); }
template <> int launch_stream2d<double, 2>(const S2Params<double, 2> &P, int dt, cudaStream_t st) { return s2_launch_it<double, 2>(P, dt, st); }
}  // namespace b2f
