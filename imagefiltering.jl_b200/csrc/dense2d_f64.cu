// dense2d_f64.cu — double (bit-exact) instantiations (kernel widths 1..32) of the dense 2-D kernel (see dense2d.cuh)
#include "dense2d.cuh"
namespace b2f {
int launch_dense2d_f64(D2Params<double> &P, int nbatch, cudaStream_t st) { return d2_dispatch<double, 1, 32>(P, nbatch, st); }
}  // namespace b2f
