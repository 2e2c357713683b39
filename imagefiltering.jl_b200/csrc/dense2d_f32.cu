// dense2d_f32.cu — float instantiations (kernel widths 1..32) of the dense 2-D kernel (see dense2d.cuh)
#include "dense2d.cuh"
namespace b2f {
int launch_dense2d_f32(D2Params<float> &P, int nbatch, cudaStream_t st) { return d2_dispatch<float, 1, 32>(P, nbatch, st); }
}  // namespace b2f
