// fft.cu — imfilter!(r::AbstractResource{FFT}, out, img, kernel, border): b2f_imfilter_fft (include/b2f.h).
//
// The reference's FFT algorithm (src/imfilter.jl:776-888): pad the image (padarray), place the kernel in a zero array of the
// padded size with periodic (FFTView) indexing, out = irfft(rfft(A) .* conj(rfft(krn))), copy the requested indices out.  The
// transforms are a LIBRARY operation there (FFTW) and here (cuFFT, loaded with dlopen on first use so that libb2f.so carries no
// link-time dependency on it); what this file owns is everything around them, as three small kernels:
//   pad_kernel      the padded image in the compute type — border remap and eltype conversion in one gather (no host padarray);
//   place_kernel    the kernel taps scattered into the zero array at their indices modulo the padded size;
//   mulconj_kernel  A_f[i] *= conj(K_f[i]) / prod(size)  (irfft's normalisation folded in);
//   crop_kernel     out[I] = filtered[I - first(padded)] for I in the output indices, converted to eltype(out).
// Arithmetic type = eltype(out) (Float32: R2C / C2R, Float64: D2Z / Z2D).  Up to 3 transformed axes; a trailing axis the kernel
// does not extend along (and the border does not pad) is a batch.  The result equals the FIR result up to the rounding of the
// transforms (the reference's tests assert `≈` between the two algorithms, test/2d.jl:69-140).
#include <cufft.h>
#include <dlfcn.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace b2f {

struct CufftApi {
    void *h = nullptr;
    cufftResult (*PlanMany)(cufftHandle *, int, int *, int *, int, int, int *, int, int, cufftType, int) = nullptr;
    cufftResult (*SetStream)(cufftHandle, cudaStream_t) = nullptr;
    cufftResult (*ExecR2C)(cufftHandle, cufftReal *, cufftComplex *) = nullptr;
    cufftResult (*ExecC2R)(cufftHandle, cufftComplex *, cufftReal *) = nullptr;
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *) = nullptr;
    cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex *, cufftDoubleReal *) = nullptr;
    cufftResult (*Destroy)(cufftHandle) = nullptr;
    bool ok = false;
};

static CufftApi &cufft_api() {
    static CufftApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    const char *names[] = {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so",
                           "libcufft.so.12", "libcufft.so.10"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.h) break;
    }
    if (!api.h) return api;
#define B2F_SYM(field, name) *(void **)(&api.field) = dlsym(api.h, name)
    B2F_SYM(PlanMany, "cufftPlanMany");
    B2F_SYM(SetStream, "cufftSetStream");
    B2F_SYM(ExecR2C, "cufftExecR2C");
    B2F_SYM(ExecC2R, "cufftExecC2R");
    B2F_SYM(ExecD2Z, "cufftExecD2Z");
    B2F_SYM(ExecZ2D, "cufftExecZ2D");
    B2F_SYM(Destroy, "cufftDestroy");
#undef B2F_SYM
    api.ok = api.PlanMany && api.SetStream && api.ExecR2C && api.ExecC2R && api.ExecD2Z && api.ExecZ2D && api.Destroy;
    return api;
}

struct FftGeom {
    int ndim;
    long long P[B2F_MAXDIM];        // padded extents
    long long n[B2F_MAXDIM];        // image extents
    long long pad_lo[B2F_MAXDIM];
    long long total;                // prod(P)
};

template <typename CT>
__global__ void fft_pad_kernel(const void *__restrict__ img, int img_dt, CT *__restrict__ dst, FftGeom G, int style, CT fill) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < G.total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i, src = 0, stride = 1;
        bool isfill = false;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            const long long p = r % G.P[d];
            r /= G.P[d];
            const long long s = remap_index(style, p - G.pad_lo[d], G.n[d]);
            if (s < 0) isfill = true;
            src += (s < 0 ? 0 : s) * stride;
            stride *= G.n[d];
        }
        dst[i] = isfill ? fill : load_elem<CT>(img, img_dt, src);
    }
}

// taps (kernel extents K, first indices klo, x fastest) -> krn[(klo + j) mod P]
template <typename CT>
__global__ void fft_place_kernel(const double *__restrict__ taps, CT *__restrict__ krn, FftGeom G, long long ntaps, long long K0, long long K1,
                                 long long K2, long long K3, long long l0, long long l1, long long l2, long long l3) {
    const long long K[4] = {K0, K1, K2, K3}, lo[4] = {l0, l1, l2, l3};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ntaps; i += (long long)gridDim.x * blockDim.x) {
        long long r = i, dst = 0, stride = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            const long long j = r % K[d];
            r /= K[d];
            long long p = (lo[d] + j) % G.P[d];
            if (p < 0) p += G.P[d];
            dst += p * stride;
            stride *= G.P[d];
        }
        // several taps can only land on one cell when the kernel is longer than the padded axis, which the padding rules out
        krn[dst] = (CT)taps[i];
    }
}

template <typename C2, typename CT>
__global__ void fft_mulconj_kernel(C2 *__restrict__ a, const C2 *__restrict__ k, long long n, CT scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const C2 x = a[i], y = k[i];
        C2 z;
        z.x = (x.x * y.x + x.y * y.y) * scale;          // x * conj(y)
        z.y = (x.y * y.x - x.x * y.y) * scale;
        a[i] = z;
    }
}

// out element I (output axes: extents O, first index ofirst relative to the padded array's first index) <- filt[I]
template <typename CT>
__global__ void fft_crop_kernel(const CT *__restrict__ filt, void *__restrict__ out, int out_dt, FftGeom G, long long ototal, long long O0,
                                long long O1, long long O2, long long O3, long long f0, long long f1, long long f2, long long f3,
                                long long r0, long long r1, long long r2, long long r3, long long e0, long long e1, long long e2, long long e3) {
    // O = extents of out; f = offset of out's first element inside the padded array; [r, e) = the requested indices (roi) per axis,
    // relative to out's first element
    const long long O[4] = {O0, O1, O2, O3}, f[4] = {f0, f1, f2, f3}, rl[4] = {r0, r1, r2, r3}, rh[4] = {e0, e1, e2, e3};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ototal; i += (long long)gridDim.x * blockDim.x) {
        long long r = i, src = 0, stride = 1;
        bool inside = true;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            const long long q = r % O[d];
            r /= O[d];
            inside = inside && q >= rl[d] && q < rh[d];
            src += (q + f[d]) * stride;
            stride *= G.P[d];
        }
        if (inside) store_elem<CT>(out, out_dt, i, filt[src]);
    }
}

template <typename CT> struct FftTypes;
template <> struct FftTypes<float> { typedef cufftComplex C; static const cufftType fwd = CUFFT_R2C, inv = CUFFT_C2R; };
template <> struct FftTypes<double> { typedef cufftDoubleComplex C; static const cufftType fwd = CUFFT_D2Z, inv = CUFFT_Z2D; };

static cufftResult exec_fwd(CufftApi &A, cufftHandle p, float *in, cufftComplex *out) { return A.ExecR2C(p, in, out); }
static cufftResult exec_fwd(CufftApi &A, cufftHandle p, double *in, cufftDoubleComplex *out) { return A.ExecD2Z(p, in, out); }
static cufftResult exec_inv(CufftApi &A, cufftHandle p, cufftComplex *in, float *out) { return A.ExecC2R(p, in, out); }
static cufftResult exec_inv(CufftApi &A, cufftHandle p, cufftDoubleComplex *in, double *out) { return A.ExecZ2D(p, in, out); }

template <typename CT>
static int run_fft_typed(const Plan &P, const b2f_array *img, const void *d_img, const b2f_array *out, void *d_out, cudaStream_t st) {
    typedef typename FftTypes<CT>::C C2;
    CufftApi &A = cufft_api();
    if (!A.ok) return fail(B2F_ENOTSUP, "cuFFT could not be loaded (libcufft.so.11): Algorithm.FFT() is unavailable");
    const StageInfo &si = P.stages[0];
    const int N = P.ndim;
    FftGeom G;
    G.ndim = N;
    G.total = 1;
    for (int d = 0; d < B2F_MAXDIM; ++d) {
        G.P[d] = P.padded_ax.len(d); G.n[d] = P.img_ax.len(d); G.pad_lo[d] = P.pad_lo[d];
        G.total *= G.P[d];
    }
    // transformed axes = up to the last axis along which the kernel extends or the border pads; the rest is a batch
    int rank = 1;
    for (int d = 0; d < N; ++d)
        if (si.lo[d] != 0 || si.hi[d] != 0 || P.pad_lo[d] != 0 || P.pad_hi[d] != 0) rank = d + 1;
    if (rank > 3) return fail(B2F_ENOTSUP, "FFT filtering transforms at most 3 axes");
    long long batch = 1, vol = 1;
    for (int d = rank; d < B2F_MAXDIM; ++d) batch *= G.P[d];
    for (int d = 0; d < rank; ++d) vol *= G.P[d];
    if (vol >= (1LL << 31) || batch >= (1LL << 31)) return fail(B2F_ENOTSUP, "array too large for the FFT path");
    const long long cvol = (G.P[0] / 2 + 1) * (vol / G.P[0]);
    for (int d = 0; d < N; ++d)
        if (si.hi[d] - si.lo[d] + 1 > G.P[d]) return fail(B2F_EDIM, "kernel longer than the padded image along axis %d", d);

    AsyncFrees guard(st);
    CT *a = nullptr, *k = nullptr;
    C2 *af = nullptr, *kf = nullptr;
    double *d_taps = nullptr;
    long long ntaps = 1;
    for (int d = 0; d < B2F_MAXDIM; ++d) ntaps *= si.hi[d] - si.lo[d] + 1;
    B2F_CUDA(cudaMallocAsync((void **)&a, sizeof(CT) * (size_t)G.total, st)); guard.push_back(a);
    B2F_CUDA(cudaMallocAsync((void **)&k, sizeof(CT) * (size_t)vol, st)); guard.push_back(k);
    B2F_CUDA(cudaMallocAsync((void **)&af, sizeof(C2) * (size_t)(cvol * batch), st)); guard.push_back(af);
    B2F_CUDA(cudaMallocAsync((void **)&kf, sizeof(C2) * (size_t)cvol, st)); guard.push_back(kf);
    B2F_CUDA(cudaMallocAsync((void **)&d_taps, sizeof(double) * (size_t)ntaps, st)); guard.push_back(d_taps);
    B2F_CUDA(cudaMemcpyAsync(d_taps, si.s->taps, sizeof(double) * (size_t)ntaps, cudaMemcpyHostToDevice, st));
    B2F_CUDA(cudaMemsetAsync(k, 0, sizeof(CT) * (size_t)vol, st));
    const int T = 256;
    auto blocks = [&](long long n) { return (unsigned)std::min<long long>((n + T - 1) / T, (long long)sm_count() * 32); };
    fft_pad_kernel<CT><<<blocks(G.total), T, 0, st>>>(d_img, img->dtype, a, G, P.style, (CT)P.fill);
    {
        FftGeom Gk = G;                      // the kernel array covers the transformed axes only
        for (int d = rank; d < B2F_MAXDIM; ++d) Gk.P[d] = 1;
        fft_place_kernel<CT><<<blocks(ntaps), T, 0, st>>>(d_taps, k, Gk, ntaps, si.hi[0] - si.lo[0] + 1, si.hi[1] - si.lo[1] + 1,
                                                          si.hi[2] - si.lo[2] + 1, si.hi[3] - si.lo[3] + 1, si.lo[0], si.lo[1], si.lo[2], si.lo[3]);
    }
    count_launch(2);
    int dims[3] = {1, 1, 1};
    for (int d = 0; d < rank; ++d) dims[d] = (int)G.P[rank - 1 - d];                // cuFFT is row-major: slowest axis first
    cufftHandle pf = 0, pk = 0, pi = 0;
    // plans are cached per thread (creating one costs milliseconds: more than the transforms of a 2048^2 image)
    struct Cached { int ty, rank, d0, d1, d2, nb, dev; cufftHandle h; };
    static thread_local std::vector<Cached> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    auto plan = [&](cufftHandle *h, cufftType ty, int nb) -> int {
        for (const Cached &c : cache)
            if (c.ty == (int)ty && c.rank == rank && c.d0 == dims[0] && c.d1 == dims[1] && c.d2 == dims[2] && c.nb == nb && c.dev == dev) {
                *h = c.h;
                return A.SetStream(*h, st) == CUFFT_SUCCESS ? 0 : fail(B2F_ECUDA, "cufftSetStream failed");
            }
        cufftResult r = A.PlanMany(h, rank, dims, nullptr, 1, 0, nullptr, 1, 0, ty, nb);
        if (r != CUFFT_SUCCESS) return fail(B2F_ECUDA, "cufftPlanMany failed (%d)", (int)r);
        r = A.SetStream(*h, st);
        if (r != CUFFT_SUCCESS) return fail(B2F_ECUDA, "cufftSetStream failed (%d)", (int)r);
        if (cache.size() >= 12) {                                    // bounded: drop the oldest plan
            cudaStreamSynchronize(st);
            A.Destroy(cache.front().h);
            cache.erase(cache.begin());
        }
        cache.push_back({(int)ty, rank, dims[0], dims[1], dims[2], nb, dev, *h});
        return 0;
    };
    int rc = plan(&pf, FftTypes<CT>::fwd, (int)batch);
    if (!rc) rc = batch == 1 ? 0 : plan(&pk, FftTypes<CT>::fwd, 1);
    if (!rc) rc = plan(&pi, FftTypes<CT>::inv, (int)batch);
    if (!rc) {
        cufftResult r = exec_fwd(A, pf, a, af);
        if (r == CUFFT_SUCCESS) r = exec_fwd(A, batch == 1 ? pf : pk, k, kf);
        if (r == CUFFT_SUCCESS) {
            for (long long b = 0; b < batch; ++b)
                fft_mulconj_kernel<C2, CT><<<blocks(cvol), T, 0, st>>>(af + b * cvol, kf, cvol, (CT)(1.0 / (double)vol));
            count_launch((int)batch);
            r = exec_inv(A, pi, af, a);
        }
        if (r != CUFFT_SUCCESS) rc = fail(B2F_ECUDA, "cuFFT execution failed (%d)", (int)r);
    }
    if (!rc) {
        long long O[4], f[4], rl[4], rh[4], ototal = 1;
        for (int d = 0; d < B2F_MAXDIM; ++d) {
            O[d] = P.out_ax.len(d);
            f[d] = P.out_ax.lo[d] - P.padded_ax.lo[d];
            rl[d] = P.roi.lo[d] - P.out_ax.lo[d];
            rh[d] = P.roi.hi[d] - P.out_ax.lo[d] + 1;
            ototal *= O[d];
        }
        fft_crop_kernel<CT><<<blocks(ototal), T, 0, st>>>(a, d_out, out->dtype, G, ototal, O[0], O[1], O[2], O[3], f[0], f[1], f[2], f[3], rl[0],
                                                          rl[1], rl[2], rl[3], rh[0], rh[1], rh[2], rh[3]);
        count_launch(1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(B2F_ECUDA, "FFT path launch failed: %s", cudaGetErrorString(e));
    }
    // the cached plans own work areas the queued transforms use: the next call on this thread may run on another stream
    cudaStreamSynchronize(st);
    return rc;
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_imfilter_fft(const b2f_array *img, const b2f_array *out, const b2f_stage *kernel, const b2f_border *border,
                                const int64_t *roi_lo, const int64_t *roi_hi, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!img || !out || !kernel || !border) return fail(B2F_EARG, "NULL argument");
    if (kernel->kind != B2F_STAGE_DENSE && kernel->kind != B2F_STAGE_1D)
        return fail(B2F_EARG, "the FFT path takes ONE array kernel (kernelconv of the factors)");
    if (out->dtype != B2F_F32 && out->dtype != B2F_F64) return fail(B2F_EINEXACT, "FFT filtering produces Float32 / Float64 arrays");
    Plan P;
    int rc = make_plan(img, out, kernel, 1, border, roi_lo, roi_hi, P);
    if (rc) return rc;
    if (P.img_ax.empty() || P.roi.empty()) { set_path("empty"); return 0; }
    // the requested indices must lie where the (valid) correlation is defined: inside the padded array shrunk by the kernel
    for (int d = 0; d < P.ndim; ++d)
        if (P.roi.lo[d] + P.stages[0].lo[d] < P.padded_ax.lo[d] || P.roi.hi[d] + P.stages[0].hi[d] > P.padded_ax.hi[d])
            return fail(B2F_EDIM, "requested indices reach outside the padded image along axis %d", d);
    rc = ensure_ctx();
    if (rc) return rc;
    set_path("fft");
    Staged sin, sout;
    rc = stage_in(img, sin, st, true);
    if (!rc) rc = stage_in(out, sout, st, roi_lo != nullptr);
    if (!rc) {
        rc = out->dtype == B2F_F32 ? run_fft_typed<float>(P, img, sin.dptr, out, sout.dptr, st) : run_fft_typed<double>(P, img, sin.dptr, out, sout.dptr, st);
        if (!rc && out->mem == B2F_HOST && sout.bytes) {
            cudaError_t e = cudaMemcpyAsync(out->ptr, sout.dptr, sout.bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) rc = fail(B2F_ECUDA, "D2H copy failed: %s", cudaGetErrorString(e));
        }
    }
    release(sin, st);
    release(sout, st);
    if (img->mem == B2F_HOST || out->mem == B2F_HOST) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess && !rc) rc = fail(B2F_ECUDA, "stream sync failed: %s", cudaGetErrorString(e));
    }
    return rc;
}
