// extrema2d_min.cu — EX_MIN instantiations of the streamed running-extrema kernel (see extrema2d.cuh)
#include "extrema2d.cuh"
namespace b2f {
int launch_extrema2d_min(const E2Params &P, cudaStream_t st) { return e2_launch<EX_MIN>(P, st); }
}  // namespace b2f
