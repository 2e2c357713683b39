// stream2d_f32_f32.cu — stream2d kernels for float images computed in float (see stream2d.cuh)
#include "stream2d_inst.cuh"
namespace b2f {
B2F_S2_INSTANTIATE(float, float)
}  // namespace b2f
