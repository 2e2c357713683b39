// stream2d_f32_f64.cu — stream2d kernels for float images computed in double (see stream2d.cuh)
#include "stream2d_inst.cuh"
namespace b2f {
B2F_S2_INSTANTIATE(float, double)
}  // namespace b2f
