// slab.cu — b2f_imfilter_slab: the per-rank piece of the slab-sharded 3-D path.
#include "common.cuh"

using namespace b2f;

extern "C" int b2f_imfilter_slab(const b2f_array *, const b2f_array *, const b2f_stage *, int32_t,
                                 const b2f_border *, int64_t, int64_t, int64_t, int64_t, void *) {
    return fail(B2F_ENOTSUP, "b2f_imfilter_slab: not built yet");
}
