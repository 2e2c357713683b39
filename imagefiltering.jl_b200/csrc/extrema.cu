// extrema.cu — K4 dispatch: running min/max/extrema (reference src/mapwindow.jl:337-481 and the
// generic path :270-333 for minimum/maximum).
#include "common.cuh"

namespace b2f {

int run_extrema_generic(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                        const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                        cudaStream_t st);

int run_extrema(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                cudaStream_t st) {
    return run_extrema_generic(img, d_img, d_min, d_max, interleaved, out_ax, wlo, whi, style, fill, st);
}

}  // namespace b2f
