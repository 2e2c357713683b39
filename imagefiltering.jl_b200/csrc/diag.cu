// diag.cu — measurement helpers of libb2f.so: the FP32 multiply-add peak of the current GPU, measured with a pure FFMA2 loop.
// bench.py reports the dense-kernel path (dense2d, BASELINE config 3) against THIS number instead of a derived one.
#include <algorithm>

#include "common.cuh"

namespace b2f {

struct DiagTaps { float k[16]; };

__global__ void __launch_bounds__(256) diag_fma_kernel(float *out, int iters, const __grid_constant__ DiagTaps p) {
    float2 acc[16];
    const float2 v = make_float2(threadIdx.x * 1e-3f, blockIdx.x * 1e-3f);
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = make_float2((float)q, (float)-q);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                unsigned long long ra = *reinterpret_cast<unsigned long long *>(&acc[q]), rc = *reinterpret_cast<const unsigned long long *>(&v), rb, rd;
                asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(p.k[j]));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
                acc[q] = *reinterpret_cast<float2 *>(&rd);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += acc[q].x + acc[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace b2f

using namespace b2f;

extern "C" int b2f_bench_fma_peak(double *tfma_per_s, void *stream) {
    if (!tfma_per_s) return fail(B2F_EARG, "NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = sm_count() * 8, iters = 4000;
    float *out = nullptr;
    B2F_CUDA(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)));
    DiagTaps p;
    for (int j = 0; j < 16; ++j) p.k[j] = 0.5f + 0.01f * j;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a, st);
        diag_fma_kernel<<<blocks, 256, 0, st>>>(out, iters, p);
        cudaEventRecord(b, st);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { cudaFree(out); return fail(B2F_ECUDA, "fma peak kernel failed: %s", cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double fma = (double)blocks * 256 * iters * 16.0 * 16.0 * 2.0;
        if (rep > 0 && ms > 0.f) best = std::max(best, fma / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tfma_per_s = best;
    return 0;
}
