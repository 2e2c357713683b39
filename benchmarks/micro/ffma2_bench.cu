// Micro-benchmark: FP32 FMA throughput on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
struct K { float k[16]; };
__device__ __forceinline__ float2 fma2s(float2 a, float k, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rc = *reinterpret_cast<unsigned long long *>(&c), rb, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(k));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
template <bool PACKED>
__global__ void __launch_bounds__(256) bench(float *out, int iters, const __grid_constant__ K p) {
    float2 acc[16];
    float2 v = make_float2(threadIdx.x * 1e-3f, blockIdx.x * 1e-3f);
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = make_float2(q, -q);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if (PACKED) acc[q] = fma2s(acc[q], p.k[j], v);
                else { acc[q].x = fmaf(acc[q].x, p.k[j], v.x); acc[q].y = fmaf(acc[q].y, p.k[j], v.y); }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += acc[q].x + acc[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    K p; for (int j = 0; j < 16; ++j) p.k[j] = 0.5f + 0.01f * j;
    float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 2000;
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            if (packed) bench<true><<<148 * 8, 256>>>(out, iters, p); else bench<false><<<148 * 8, 256>>>(out, iters, p);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double fma = 148.0 * 8 * 256 * (double)iters * 16 * 16 * 2;
            printf("%s: %.3f ms  %.2f TFMA/s\n", packed ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9);
        }
    }
    return 0;
}
