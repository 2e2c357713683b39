// stream2d_f64_f32.cu — stream2d kernels for double images computed in float (see stream2d.cuh)
#include "stream2d_inst.cuh"
namespace b2f {
B2F_S2_INSTANTIATE(double, float)
}  // namespace b2f
