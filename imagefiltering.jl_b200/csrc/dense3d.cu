// dense3d.cu — one dense (non-separable) 3-D stage: the reference's dense loop (src/imfilter.jl:624-669) for N = 3, e.g. the
// 3 x 3 x 3 "densesmall" and the 13 x 13 x 13 DoG "denselarge" kernels of its benchmark suite on 100^3 volumes
// (benchmark/benchmarks.jl:36-49).  Before this kernel those ran on the per-tap generic path (13^3 on 100^3: 13.7 ms).
//
// A CTA (256 threads) owns 32 x 8 x 4 outputs.  The input block with its halo ((32+Kx-1) x (8+Ky-1) x (4+Kz-1), border remap
// and eltype conversion applied once) and the taps sit in shared memory; a thread owns 4 adjacent outputs along x and walks
// the taps in the reference's order (x fastest, then y, then z — Float64 results are bit-equal to the CPU loop: separate
// multiply and add), reading its row window with 128-bit loads: 4 x Kx multiply-adds per (4 + Kx - 1) values.  A 4-th axis is
// a batch.  Bound: the FP32 / FP64 pipe (Kx Ky Kz multiply-adds per voxel against 4 + 4 or 4 + 8 bytes).
#include "common.cuh"

namespace b2f {

constexpr int D3_TX = 32, D3_TY = 8, D3_TZ = 4, D3_NT = 256;
constexpr size_t D3_SMEM_MAX = 200 * 1024;

template <typename CT>
struct D3Params {
    const void *img;
    CT *out;
    const CT *taps;              // device, Kx * Ky * Kz, x fastest
    int img_dt, style;
    int W, H, D;                 // image extents along x, y, z
    long long img_vol;           // W * H * D (batch stride)
    int Kx, Ky, Kz, klox, kloy, kloz;
    int rx0, ry0, rz0, rw, rh, rd;          // outputs: image positions [r*0, r*0 + r*) per axis
    int ox, oy, oz;                         // position of out's first element in image coordinates
    long long opx, opy, ovol;               // out pitches: row, plane, volume (batch stride)
    int ntx, nty, ntz;
    int pitch, rows, planes;                // tile geometry in shared memory
    CT fill;
};

static __device__ __noinline__ int d3_remap_slow(int style, int i, int n) { return (int)remap_index(style, (int64_t)i, (int64_t)n); }
__device__ __forceinline__ int d3_remap(int style, int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    return d3_remap_slow(style, i, n);
}

// four consecutive tile values, 16-byte aligned (the pitch is a multiple of 4 and a thread starts at 4 * tx)
__device__ __forceinline__ void d3_load4(const float *p, float *w) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
}
__device__ __forceinline__ void d3_load4(const double *p, double *w) {
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

template <typename CT>
__global__ void __launch_bounds__(D3_NT) dense3d_kernel(const __grid_constant__ D3Params<CT> P) {
    extern __shared__ __align__(16) unsigned char d3_smem[];
    CT *tile = reinterpret_cast<CT *>(d3_smem);
    CT *taps = tile + (size_t)P.pitch * P.rows * P.planes;
    const int tid = threadIdx.x;
    long long t = blockIdx.x;
    const int bx = (int)(t % P.ntx); t /= P.ntx;
    const int by = (int)(t % P.nty); t /= P.nty;
    const int bz = (int)(t % P.ntz);
    const long long b = t / P.ntz;
    const int x0 = P.rx0 + bx * D3_TX, y0 = P.ry0 + by * D3_TY, z0 = P.rz0 + bz * D3_TZ;     // first output of the tile
    const int ntap = P.Kx * P.Ky * P.Kz;
    for (int i = tid; i < ntap; i += D3_NT) taps[i] = P.taps[i];
    // tile load: lanes along x; cells past the halo (pitch padding) are zero
    const int ncell = P.pitch * P.rows * P.planes, incols = D3_TX + P.Kx - 1;
    const char *base = reinterpret_cast<const char *>(P.img);
    for (int i = tid; i < ncell; i += D3_NT) {
        const int c = i % P.pitch, r = (i / P.pitch) % P.rows, p = i / (P.pitch * P.rows);
        CT v = (CT)0;
        if (c < incols) {
            const int sx = d3_remap(P.style, x0 + P.klox + c, P.W), sy = d3_remap(P.style, y0 + P.kloy + r, P.H),
                      sz = d3_remap(P.style, z0 + P.kloz + p, P.D);
            v = (sx < 0 || sy < 0 || sz < 0) ? P.fill : load_elem<CT>(base, P.img_dt, b * P.img_vol + ((long long)sz * P.H + sy) * P.W + sx);
        }
        tile[i] = v;
    }
    __syncthreads();
    const int tx = tid & 7, ty = (tid >> 3) & 7, tz = tid >> 6;
    CT acc[4] = {(CT)0, (CT)0, (CT)0, (CT)0};
    const CT *tp = taps;
    for (int kz = 0; kz < P.Kz; ++kz)
        for (int ky = 0; ky < P.Ky; ++ky) {
            const CT *row = tile + ((size_t)(tz + kz) * P.rows + (ty + ky)) * P.pitch + 4 * tx;
            CT w[8];
            d3_load4(row, w);
            for (int kx0 = 0; kx0 < P.Kx; kx0 += 4) {
                d3_load4(row + kx0 + 4, w + 4);                                 // inside the padded pitch
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (kx0 + kk < P.Kx) {
                        const CT k = tp[kx0 + kk];
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[i] = mac<CT>(acc[i], w[i + kk], k);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = w[4 + i];
            }
            tp += P.Kx;
        }
    const int gx = x0 + 4 * tx, gy = y0 + ty, gz = z0 + tz;
    if (gy < P.ry0 + P.rh && gz < P.rz0 + P.rd) {
        CT *o = P.out + b * P.ovol + (long long)(gz - P.oz) * P.opy + (long long)(gy - P.oy) * P.opx + (gx - P.ox);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (gx + i < P.rx0 + P.rw) o[i] = acc[i];
    }
}

static bool d3_geometry(int64_t Kx, int64_t Ky, int64_t Kz, size_t esz, int &pitch, int &rows, int &planes, size_t &smem) {
    pitch = (int)(((D3_TX + Kx - 1 + 4) + 3) / 4 * 4);           // + 4: the window runs one 4-group past the last tap group
    rows = (int)(D3_TY + Ky - 1);
    planes = (int)(D3_TZ + Kz - 1);
    smem = ((size_t)pitch * rows * planes + (size_t)(Kx * Ky * Kz)) * esz;
    return smem <= D3_SMEM_MAX;
}

// exactly one active stage, dense, with extent along axis 2 (otherwise dense2d is the kernel); Float32 / Float64 output
bool dense3d_applicable(const Plan &P, int img_dt, int out_dt) {
    if (P.ndim < 3 || P.active.size() != 1) return false;
    if (out_dt != B2F_F32 && out_dt != B2F_F64) return false;
    const StageInfo &si = P.stages[P.active[0]];
    if (si.s->kind != B2F_STAGE_DENSE) return false;
    if (si.lo[3] != 0 || si.hi[3] != 0) return false;
    if (P.roi.lo[3] != P.img_ax.lo[3] || P.roi.hi[3] != P.img_ax.hi[3] || P.out_ax.lo[3] != P.img_ax.lo[3] || P.out_ax.hi[3] != P.img_ax.hi[3])
        return false;
    const int64_t Kx = si.hi[0] - si.lo[0] + 1, Ky = si.hi[1] - si.lo[1] + 1, Kz = si.hi[2] - si.lo[2] + 1;
    if (Kz < 2 || Kx > 64 || Ky > 64 || Kz > 64) return false;
    for (int d = 0; d < 3; ++d)
        if (P.img_ax.len(d) >= (1LL << 30)) return false;
    int pitch, rows, planes;
    size_t smem;
    (void)img_dt;
    return d3_geometry(Kx, Ky, Kz, out_dt == B2F_F32 ? 4 : 8, pitch, rows, planes, smem);
}

template <typename CT>
static int run_dense3d_typed(const Plan &P0, const void *d_img, int img_dt, void *d_out, cudaStream_t st) {
    const StageInfo &si = P0.stages[P0.active[0]];
    D3Params<CT> P;
    memset(&P, 0, sizeof P);
    P.img = d_img; P.img_dt = img_dt; P.out = (CT *)d_out; P.style = P0.style; P.fill = (CT)P0.fill;
    P.W = (int)P0.img_ax.len(0); P.H = (int)P0.img_ax.len(1); P.D = (int)P0.img_ax.len(2);
    P.img_vol = (long long)P.W * P.H * P.D;
    P.Kx = (int)(si.hi[0] - si.lo[0] + 1); P.Ky = (int)(si.hi[1] - si.lo[1] + 1); P.Kz = (int)(si.hi[2] - si.lo[2] + 1);
    P.klox = (int)si.lo[0]; P.kloy = (int)si.lo[1]; P.kloz = (int)si.lo[2];
    P.rx0 = (int)(P0.roi.lo[0] - P0.img_ax.lo[0]); P.ry0 = (int)(P0.roi.lo[1] - P0.img_ax.lo[1]); P.rz0 = (int)(P0.roi.lo[2] - P0.img_ax.lo[2]);
    P.rw = (int)P0.roi.len(0); P.rh = (int)P0.roi.len(1); P.rd = (int)P0.roi.len(2);
    P.ox = (int)(P0.out_ax.lo[0] - P0.img_ax.lo[0]); P.oy = (int)(P0.out_ax.lo[1] - P0.img_ax.lo[1]); P.oz = (int)(P0.out_ax.lo[2] - P0.img_ax.lo[2]);
    P.opx = P0.out_ax.len(0); P.opy = P.opx * P0.out_ax.len(1); P.ovol = P.opy * P0.out_ax.len(2);
    P.ntx = (P.rw + D3_TX - 1) / D3_TX; P.nty = (P.rh + D3_TY - 1) / D3_TY; P.ntz = (P.rd + D3_TZ - 1) / D3_TZ;
    size_t smem;
    if (!d3_geometry(P.Kx, P.Ky, P.Kz, sizeof(CT), P.pitch, P.rows, P.planes, smem)) return fail(B2F_ENOTSUP, "dense3d: kernel too large");
    const long long nbatch = P0.img_ax.len(3);
    const long long blocks = (long long)P.ntx * P.nty * P.ntz * nbatch;
    if (blocks <= 0) return 0;
    if (blocks >= (1LL << 31)) return fail(B2F_ENOTSUP, "dense3d: array too large for one launch");
    const size_t ntap = (size_t)P.Kx * P.Ky * P.Kz;
    std::vector<CT> h(ntap);
    for (size_t i = 0; i < ntap; ++i) h[i] = (CT)si.s->taps[i];
    CT *d_taps = nullptr;
    B2F_CUDA(cudaMallocAsync((void **)&d_taps, ntap * sizeof(CT), st));
    AsyncFrees guard(st);
    guard.push_back(d_taps);
    B2F_CUDA(cudaMemcpyAsync(d_taps, h.data(), ntap * sizeof(CT), cudaMemcpyHostToDevice, st));
    P.taps = d_taps;
    if (smem > 48 * 1024)                        // per launch: the attribute belongs to the current device's copy of the kernel
        B2F_CUDA(cudaFuncSetAttribute(dense3d_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D3_SMEM_MAX));
    dense3d_kernel<CT><<<(unsigned)blocks, D3_NT, smem, st>>>(P);
    count_launch(1);
    B2F_CUDA(cudaGetLastError());
    return 0;
}

int run_dense3d(const Plan &P, const void *d_img, int img_dt, void *d_out, int out_dt, cudaStream_t st) {
    set_path("dense3d");
    return out_dt == B2F_F32 ? run_dense3d_typed<float>(P, d_img, img_dt, d_out, st) : run_dense3d_typed<double>(P, d_img, img_dt, d_out, st);
}

}  // namespace b2f
