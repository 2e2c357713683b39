// extrema.cu — K4 dispatch: running min/max/extrema (reference src/mapwindow.jl:337-481 and the
// generic path :270-333 for minimum/maximum).
#include "extrema2d.cuh"

namespace b2f {

int run_extrema_generic(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                        const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                        cudaStream_t st);
// van Herk / Gil-Werman running extrema: any window width, any eltype, any axis (extrema_vh.cu)
bool extrema_vh_applicable(const b2f_array *img, const int64_t *wlo, const int64_t *whi);
int run_extrema_vh(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved, const Box &out_ax,
                   const int64_t *wlo, const int64_t *whi, int style, double fill, cudaStream_t st);

// Float32 images, window on axes 0/1 only (later axes are a batch), window contains its centre, <= 16 wide.
static bool extrema2d_applicable(const b2f_array *img, const Box &out_ax, const int64_t *wlo, const int64_t *whi) {
    if (img->dtype != B2F_F32 || img->ndim < 1) return false;
    Box ia = axes_of(img);
    for (int d = 0; d < img->ndim; ++d) {
        if (d >= 2) {
            if (wlo[d] != 0 || whi[d] != 0) return false;
            if (out_ax.lo[d] != ia.lo[d] || out_ax.hi[d] != ia.hi[d]) return false;
        } else {
            if (wlo[d] > 0 || whi[d] < 0 || whi[d] - wlo[d] + 1 > 16) return false;
        }
    }
    if (ia.len(0) >= (1LL << 30) || ia.len(1) >= (1LL << 30)) return false;
    return true;
}

int run_extrema(const b2f_array *img, const void *d_img, void *d_min, void *d_max, int interleaved,
                const Box &out_ax, const int64_t *wlo, const int64_t *whi, int style, double fill,
                cudaStream_t st) {
    const char *force = getenv("B2F_FORCE_PATH");     // debugging knob: "generic" | "vh" bypass the faster kernels
    const bool force_generic = force && strcmp(force, "vh") != 0;
    if (force || !extrema2d_applicable(img, out_ax, wlo, whi)) {
        if (!force_generic && extrema_vh_applicable(img, wlo, whi)) {
            int rc = run_extrema_vh(img, d_img, d_min, d_max, interleaved, out_ax, wlo, whi, style, fill, st);
            if (rc != B2F_ENOTSUP) return rc;
        }
        return run_extrema_generic(img, d_img, d_min, d_max, interleaved, out_ax, wlo, whi, style, fill, st);
    }
    set_path("extrema2d");
    Box ia = axes_of(img);
    E2Params P;
    memset(&P, 0, sizeof P);
    P.img = d_img;
    P.W = (int)ia.len(0); P.H = (int)ia.len(1);
    P.img_plane = (long long)P.W * P.H;
    P.omin = d_min; P.omax = d_max;
    P.out_pitch = out_ax.len(0);
    P.out_plane = out_ax.len(0) * out_ax.len(1);
    P.out_ox = (int)(out_ax.lo[0] - ia.lo[0]); P.out_oy = (int)(out_ax.lo[1] - ia.lo[1]);
    P.rx0 = P.out_ox; P.ry0 = P.out_oy; P.rw = (int)out_ax.len(0); P.rh = (int)out_ax.len(1);
    P.style = style == B2F_FILL ? B2F_FILL : B2F_REPLICATE;   // truncation == replicate for min/max
    P.fill = (float)fill;
    P.Wx = (int)(whi[0] - wlo[0] + 1); P.lox = (int)wlo[0];
    P.Wy = img->ndim > 1 ? (int)(whi[1] - wlo[1] + 1) : 1; P.loy = img->ndim > 1 ? (int)wlo[1] : 0;
    const int PXo = 4;
    bool aligned = (P.out_pitch % PXo == 0) && (P.out_plane % PXo == 0);
    if (d_min) aligned = aligned && reinterpret_cast<uintptr_t>(d_min) % 16 == 0;
    if (d_max) aligned = aligned && reinterpret_cast<uintptr_t>(d_max) % 16 == 0;
    P.vec_ok = aligned;
    const long long nbatch = ia.len(2) * ia.len(3);
    P.nsx = (P.rw + 127) / 128;
    const long long want = (long long)sm_count() * 16 * 6;
    int SH = 256;
    while (SH > 32 && (long long)P.nsx * ((P.rh + SH - 1) / SH) * nbatch < want) SH >>= 1;
    P.SH = SH;
    P.nsy = (P.rh + SH - 1) / SH;
    P.nstrips = (long long)P.nsx * P.nsy * nbatch;
    if (interleaved) return launch_extrema2d_pair(P, st);
    if (d_min && d_max) return launch_extrema2d_both(P, st);
    return d_min ? launch_extrema2d_min(P, st) : launch_extrema2d_max(P, st);
}

}  // namespace b2f
