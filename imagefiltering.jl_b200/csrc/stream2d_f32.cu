// stream2d_f32.cu — float instantiations of the warp-streamed fused 2-D separable kernel
#include "stream2d_inst.cuh"
namespace b2f {
template <> int launch_stream2d<float, 1>(const S2Params<float, 1> &P, int dt, cudaStream_t st) { return s2_launch_it<float, 1>(P, dt, st); }
template <> int launch_stream2d<float, 2>(const S2Params<float, 2> &P, int dt, cudaStream_t st) { return s2_launch_it<float, 2>(P, dt, st); }
}  // namespace b2f
