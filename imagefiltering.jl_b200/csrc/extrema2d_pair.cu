// extrema2d_pair.cu — EX_PAIR instantiations of the streamed running-extrema kernel (see extrema2d.cuh)
#include "extrema2d.cuh"
namespace b2f {
int launch_extrema2d_pair(const E2Params &P, cudaStream_t st) { return e2_launch<EX_PAIR>(P, st); }
}  // namespace b2f
