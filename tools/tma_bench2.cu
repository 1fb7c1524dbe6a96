// Microbenchmark 2: does alternating between two tensor maps (activation / weight) or issuing from two
// warps change the UTMALDG issue cost?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench2 tma_bench2.cu -lcuda
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
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
constexpr int NL = 6;
// mode 0: one thread issues A,A,A,...   mode 1: one thread alternates A,B   mode 2: warp0 issues A's, warp1 issues B's
// mode 3: one thread, A with out-of-bounds coordinates (-1 offsets)  mode 4: like 1 but prefetch.tensormap of the other map before each use
__global__ void bench(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int mode, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2 * NL];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * NL; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long* o = out + (size_t)blockIdx.x * 64;
    const int h0 = (blockIdx.x * 4) % 150;
    if (lane == 0 && warp == 0) {
        const long long t0 = clock64();
        for (int i = 0; i < NL; ++i) {
            const uint32_t b = smem_u32(&bar[i]);
            if (mode == 1 || mode == 4) {
                if (i & 1) { mbar_expect_tx(b, 8192); tma2d(base + i * 16384, &tb, b, (i * 64) % 512, 0); }
                else { mbar_expect_tx(b, 16384); tma5d(base + i * 16384, &ta, b, 0, 0, 0, h0 + i, 0); }
                if (mode == 4) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)((i & 1) ? &ta : &tb)) : "memory");
            } else if (mode == 3) {
                mbar_expect_tx(b, 16384); tma5d(base + i * 16384, &ta, b, 0, -1, 0, h0 + i - 1, 0);
            } else {
                mbar_expect_tx(b, 16384); tma5d(base + i * 16384, &ta, b, 0, 0, 0, h0 + i, 0);
            }
            o[i] = clock64() - t0;
        }
        for (int i = 0; i < NL; ++i) { mbar_wait(smem_u32(&bar[i]), 0); o[8 + i] = clock64() - t0; }
    }
    if (lane == 0 && warp == 1 && mode == 2) {
        const long long t0 = clock64();
        for (int i = 0; i < NL; ++i) {
            const uint32_t b = smem_u32(&bar[NL + i]);
            mbar_expect_tx(b, 8192); tma2d(base + (NL + i) * 16384, &tb, b, (i * 64) % 512, 0);
            o[16 + i] = clock64() - t0;
        }
        for (int i = 0; i < NL; ++i) { mbar_wait(smem_u32(&bar[NL + i]), 0); o[24 + i] = clock64() - t0; }
    }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const int H = 160, W = 160, C = 64;
    __half *buf, *wts;
    CK(cudaMalloc(&buf, (size_t)H * W * C * 2)); CK(cudaMemset(buf, 0, (size_t)H * W * C * 2));
    CK(cudaMalloc(&wts, (size_t)64 * 576 * 2)); CK(cudaMemset(wts, 0, (size_t)64 * 576 * 2));
    CUtensorMap ta, tb; cuuint32_t es[5] = {1, 1, 1, 1, 1};
    { cuuint64_t dims[5] = {C, W, 1, H, 1}; cuuint64_t str[4] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[5] = {64, 32, 1, 4, 1};
      if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc a failed\n"); return 1; } }
    { cuuint64_t dims[2] = {576, 64}; cuuint64_t str[1] = {576 * 2}; cuuint32_t box[2] = {64, 64};
      if (enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, wts, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc b failed\n"); return 1; } }
    const int smem = 2 * NL * 16384 + 1024;
    CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const char* names[5] = {"one thread: A only", "one thread: A,B alternating", "two warps: A | B", "one thread: A with OOB (-1,-1)", "one thread: A,B alternating + prefetch.tensormap"};
    for (int grid : {1, 148}) for (int mode = 0; mode < 5; ++mode) {
        long long* d; CK(cudaMalloc(&d, sizeof(long long) * 64 * grid)); CK(cudaMemset(d, 0, sizeof(long long) * 64 * grid));
        for (int r = 0; r < 3; ++r) bench<<<grid, 64, smem>>>(ta, tb, mode, d);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(64 * grid); CK(cudaMemcpy(h.data(), d, sizeof(long long) * 64 * grid, cudaMemcpyDeviceToHost));
        auto med = [&](int s) { std::vector<long long> v; for (int b = 0; b < grid; ++b) v.push_back(h[b * 64 + s]); std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
        printf("%-50s grid %3d | issue:", names[mode], grid); for (int i = 0; i < NL; ++i) printf(" %5lld", med(i));
        printf(" | done:"); for (int i = 0; i < NL; ++i) printf(" %5lld", med(8 + i));
        if (mode == 2) { printf(" || warp1 issue:"); for (int i = 0; i < NL; ++i) printf(" %5lld", med(16 + i)); printf(" | done:"); for (int i = 0; i < NL; ++i) printf(" %5lld", med(24 + i)); }
        printf("\n"); cudaFree(d);
    }
    return 0;
}
