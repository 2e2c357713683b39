// stream2d_u8_f32.cu — stream2d kernels for uint8_t images computed in float (see stream2d.cuh)
#include "stream2d_inst.cuh"
namespace b2f {
B2F_S2_INSTANTIATE(uint8_t, float)
}  // namespace b2f
