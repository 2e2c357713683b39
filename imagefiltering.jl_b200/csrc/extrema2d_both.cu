// extrema2d_both.cu — EX_BOTH instantiations of the streamed running-extrema kernel (see extrema2d.cuh)
#include "extrema2d.cuh"
namespace b2f {
int launch_extrema2d_both(const E2Params &P, cudaStream_t st) { return e2_launch<EX_BOTH>(P, st); }
}  // namespace b2f
