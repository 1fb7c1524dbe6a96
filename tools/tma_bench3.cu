// Microbenchmark 3: TMA *processing* throughput per SM for 16 KB boxes of different rank / shape.
// Six warps each issue one load at the same time (issue cost out of the picture); the time from the
// first issue to the last completion gives bytes/clk per SM.  Run warm (second pass over the same tiles).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
template <int RANK>
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    if constexpr (RANK == 2)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
    else if constexpr (RANK == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else if constexpr (RANK == 4)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
constexpr int NW = 6;
// hsel: which coordinate carries the tile's row offset; tiles are distinct per (block, warp, rep)
template <int RANK>
__global__ void bench(const __grid_constant__ CUtensorMap tm, int bytes, int hdim, int hstep, int hmax, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[NW];
    __shared__ long long t_first, t_last;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NW; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (lane == 0) {
            int c[5] = {0, 0, 0, 0, 0};
            c[hdim] = ((blockIdx.x * NW + warp) * hstep) % hmax;
            mbar_expect_tx(smem_u32(&bar[warp]), bytes);
            tma_load<RANK>(base + warp * 16384, &tm, smem_u32(&bar[warp]), c[0], c[1], c[2], c[3], c[4]);
        }
        __syncwarp();
        for (int i = 0; i < NW; ++i) mbar_wait(smem_u32(&bar[i]), rep & 1);
        const long long t1 = clock64();
        __syncthreads();
        if (rep > 0 && threadIdx.x == 0) acc += t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc / (reps - 1);
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int RANK>
void run(const char* name, EncodeFn enc, void* buf, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
         int hdim, int hstep, int hmax, int grid) {
    CUtensorMap tm; cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, RANK, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    int bytes = 2; for (int i = 0; i < RANK; ++i) bytes *= box[i];
    long long* d; CK(cudaMalloc(&d, sizeof(long long) * grid));
    const int smem = NW * 16384 + 1024;
    CK(cudaFuncSetAttribute(bench<RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    bench<RANK><<<grid, NW * 32, smem>>>(tm, bytes, hdim, hstep, hmax, 9, d);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid); CK(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    std::sort(h.begin(), h.end());
    const double med = (double)h[grid / 2];
    printf("%-46s grid %3d: %d x %5d B in %6.0f clk (median CTA) = %6.1f B/clk/SM  [min %lld max %lld]\n", name, grid, NW, bytes, med,
           NW * bytes / med, h[0], h[grid - 1]);
    cudaFree(d);
}
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const int H = 640, W = 160, C = 64;     // 13 MB activation: L2 resident, distinct tiles per warp / CTA
    __half* buf; CK(cudaMalloc(&buf, (size_t)H * W * C * 2)); CK(cudaMemset(buf, 0, (size_t)H * W * C * 2));
    for (int grid : {1, 16, 148}) {
        { cuuint64_t dims[2] = {C, (cuuint64_t)H * W}; cuuint64_t str[1] = {C * 2}; cuuint32_t box[2] = {64, 128};
          run<2>("2D {64,NHW} box {64,128}", enc, buf, dims, str, box, 1, 128, H * W - 128, grid); }
        { cuuint64_t dims[3] = {C, W, H}; cuuint64_t str[2] = {C * 2, (cuuint64_t)W * C * 2};
          cuuint32_t b1[3] = {64, 32, 4}; run<3>("3D {64,W,H} box {64,32,4}", enc, buf, dims, str, b1, 2, 4, H - 4, grid);
          cuuint32_t b2[3] = {64, 128, 1}; run<3>("3D {64,W,H} box {64,128,1}", enc, buf, dims, str, b2, 2, 1, H - 1, grid);
          cuuint32_t b3[3] = {64, 8, 16}; run<3>("3D {64,W,H} box {64,8,16}", enc, buf, dims, str, b3, 2, 16, H - 16, grid);
          cuuint32_t b4[3] = {64, 16, 8}; run<3>("3D {64,W,H} box {64,16,8}", enc, buf, dims, str, b4, 2, 8, H - 8, grid); }
        { cuuint64_t dims[4] = {C, W, H, 1}; cuuint64_t str[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
          cuuint32_t box[4] = {64, 32, 4, 1}; run<4>("4D {64,W,H,1} box {64,32,4,1}", enc, buf, dims, str, box, 2, 4, H - 4, grid); }
        { cuuint64_t dims[5] = {C, W, 1, H, 1}; cuuint64_t str[4] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
          cuuint32_t box[5] = {64, 32, 1, 4, 1}; run<5>("5D {64,W,1,H,1} box {64,32,1,4,1}", enc, buf, dims, str, box, 3, 4, H - 4, grid); }
        { // channel pitch 128 (view of a concat buffer): rows 256 B apart
          cuuint64_t dims[2] = {C, (cuuint64_t)H * W / 2}; cuuint64_t str[1] = {C * 4}; cuuint32_t box[2] = {64, 128};
          run<2>("2D pitch 256 B box {64,128}", enc, buf, dims, str, box, 1, 128, H * W / 2 - 128, grid); }
        { // 32-channel inner (64 B rows, would be SW64 in the conv): box {32, 256} = 16 KB
          cuuint64_t dims[2] = {32, (cuuint64_t)H * W * 2}; cuuint64_t str[1] = {64}; cuuint32_t box[2] = {32, 256};
          run<2>("2D {32,..} box {32,256} (64 B rows, SW128 enc)", enc, buf, dims, str, box, 1, 256, H * W * 2 - 256, grid); }
    }
    return 0;
}
