// Microbenchmark: issue rate / latency / throughput of tiled-mode TMA loads for the box shapes the
// implicit-GEMM conv uses.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bench tma_bench.cu
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
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* m, uint32_t bar, const int* c) {
    if constexpr (RANK == 2)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c[0]), "r"(c[1]) : "memory");
    else if constexpr (RANK == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]) : "memory");
    else if constexpr (RANK == 4)
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]) : "memory");
}

constexpr int NLOAD = 8;

// one thread per CTA issues NLOAD box loads back to back (distinct smem slots), then waits in order
template <int RANK>
__global__ void bench(const __grid_constant__ CUtensorMap tm, int bytes, int4 step, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[NLOAD];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NLOAD; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        long long* o = out + (size_t)blockIdx.x * 32;
        const long long t0 = clock64();
        for (int i = 0; i < NLOAD; ++i) {
            // coordinates: tile index varies with block and i so that different rows are touched
            int c[5] = {0, 0, 0, 0, 0};
            const int tile = blockIdx.x * NLOAD + i;
            if (RANK == 2) { c[1] = (tile * step.x) % step.w; }
            else if (RANK == 3) { c[1] = 0; c[2] = (tile * step.x) % step.w; }
            else if (RANK == 4) { c[1] = 0; c[2] = (tile * step.x) % step.w; c[3] = 0; }
            else { c[1] = 0; c[2] = 0; c[3] = (tile * step.x) % step.w; c[4] = 0; }
            mbar_expect_tx(smem_u32(&bar[i]), bytes);
            tma_load<RANK>(base + i * 16384, &tm, smem_u32(&bar[i]), c);
            o[i] = clock64() - t0;
        }
        for (int i = 0; i < NLOAD; ++i) {
            mbar_wait(smem_u32(&bar[i]), 0);
            o[8 + i] = clock64() - t0;
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int RANK>
void run(const char* name, EncodeFn enc, void* buf, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
         int4 step, int grid, CUtensorMapL2promotion prom = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    CUtensorMap tm;
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, RANK, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    int bytes = 2;
    for (int i = 0; i < RANK; ++i) bytes *= box[i];
    long long* d_out;
    CK(cudaMalloc(&d_out, sizeof(long long) * 32 * grid));
    CK(cudaFuncSetAttribute(bench<RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, NLOAD * 16384 + 1024));
    for (int rep = 0; rep < 3; ++rep) bench<RANK><<<grid, 32, NLOAD * 16384 + 1024>>>(tm, bytes, step, d_out);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(32 * grid);
    CK(cudaMemcpy(h.data(), d_out, sizeof(long long) * 32 * grid, cudaMemcpyDeviceToHost));
    // median over CTAs
    auto med = [&](int slot) { std::vector<long long> v; for (int b = 0; b < grid; ++b) v.push_back(h[b * 32 + slot]); std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    printf("%-44s grid %3d bytes %5d | issue:", name, grid, bytes);
    for (int i = 0; i < NLOAD; ++i) printf(" %5lld", med(i));
    printf(" | done:");
    for (int i = 0; i < NLOAD; ++i) printf(" %5lld", med(8 + i));
    const double bpc = (double)bytes * (NLOAD - 1) / (double)(med(8 + NLOAD - 1) - med(8));
    printf(" | steady %.1f B/clk/SM\n", bpc);
    cudaFree(d_out);
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    const int H = 160, W = 160, C = 64;
    __half* buf;
    CK(cudaMalloc(&buf, (size_t)H * W * C * 2 * 4));
    CK(cudaMemset(buf, 0, (size_t)H * W * C * 2 * 4));
    for (int grid : {1, 148, 296}) {
        {   // 2D: [C][rows], box 64 x 128 rows (GEMM-like)
            cuuint64_t dims[2] = {C, (cuuint64_t)H * W}; cuuint64_t str[1] = {C * 2}; cuuint32_t box[2] = {64, 128};
            run<2>("2D {64,25600} box {64,128}", enc, buf, dims, str, box, make_int4(128, 0, 0, H * W - 128), grid);
            run<2>("2D same, L2 promotion 128B", enc, buf, dims, str, box, make_int4(128, 0, 0, H * W - 128), grid, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
            run<2>("2D same, no L2 promotion", enc, buf, dims, str, box, make_int4(128, 0, 0, H * W - 128), grid, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        }
        {   // 3D: [C][W][H], box 64 x 32 x 4
            cuuint64_t dims[3] = {C, W, H}; cuuint64_t str[2] = {C * 2, (cuuint64_t)W * C * 2}; cuuint32_t box[3] = {64, 32, 4};
            run<3>("3D {64,160,160} box {64,32,4}", enc, buf, dims, str, box, make_int4(4, 0, 0, H - 4), grid);
            cuuint32_t box2[3] = {64, 128, 1};
            run<3>("3D {64,160,160} box {64,128,1}", enc, buf, dims, str, box2, make_int4(1, 0, 0, H - 1), grid);
            cuuint32_t box3[3] = {64, 8, 16};
            run<3>("3D {64,160,160} box {64,8,16}", enc, buf, dims, str, box3, make_int4(16, 0, 0, H - 16), grid);
        }
        {   // 4D: [C][W][H][N], box 64 x 32 x 4 x 1
            cuuint64_t dims[4] = {C, W, H, 1}; cuuint64_t str[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2}; cuuint32_t box[4] = {64, 32, 4, 1};
            run<4>("4D {64,160,160,1} box {64,32,4,1}", enc, buf, dims, str, box, make_int4(4, 0, 0, H - 4), grid);
        }
        {   // 5D as in conv.cu (stride 1): [C][W][1][H][N]
            cuuint64_t dims[5] = {C, W, 1, H, 1};
            cuuint64_t str[4] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
            cuuint32_t box[5] = {64, 32, 1, 4, 1};
            run<5>("5D {64,160,1,160,1} box {64,32,1,4,1}", enc, buf, dims, str, box, make_int4(4, 0, 0, H - 4), grid);
        }
        {   // 2D with 128-channel pitch (concat buffer view): rows are 256 B apart
            cuuint64_t dims[2] = {C, (cuuint64_t)H * W}; cuuint64_t str[1] = {C * 4}; cuuint32_t box[2] = {64, 128};
            run<2>("2D pitch 128ch box {64,128}", enc, buf, dims, str, box, make_int4(128, 0, 0, H * W - 128), grid);
            cuuint32_t boxb[2] = {64, 64};
            run<2>("2D box {64,64} (8 KB)", enc, buf, dims, str, boxb, make_int4(64, 0, 0, H * W - 64), grid);
        }
    }
    return 0;
}
