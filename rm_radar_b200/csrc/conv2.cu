// Round-2 implicit-GEMM convolution: one persistent CTA per SM, the whole shared memory as operand
// storage, double-buffered TMEM accumulators.  Same GEMM view and operand layouts as conv.cu
// (D[128 pixels, N] += A[128, 16] * W[N, 16]^T per tcgen05.mma, A staged by TMA straight from the NHWC
// activation, zero fill = padding); what changed is everything around the MMA:
//
//   * r1 ran 2-3 short-lived CTAs per SM with a 3-4 stage ring each.  One ring round trip (commit ->
//     empty barrier -> producer wake-up -> TMA issue -> L2 latency -> full barrier) is ~1400 cycles, so a
//     4-stage ring cannot deliver a k-block faster than every ~350-480 cycles whatever the MMA costs —
//     exactly the period profiles/r1_timeline_*.txt shows.  Here the ring owns up to 223 KB.
//   * an SS-mode MMA reads A (4 KB) and B (N x 32 B) from shared memory for every K = 16 slice and the
//     shared-memory port moves 128 B/clk: the MMA rate is max(N/2, (4096 + 32 N)/128) cycles, and TMA
//     writes compete for the same port (tools/umma_bench.cu).  So the tile takes all of Cout it can
//     (N up to 256), the weight slice of a CTA stays resident in shared memory across its tiles
//     whenever it fits (every tile after the first then streams activations only), and 3x3 stride-1
//     layers read one halo patch per 64 channels instead of nine shifted boxes.
//   * a CTA walks over its tiles (fixed output-channel slice, pixel tiles j, j + G, ...): barrier
//     init, TMEM allocation, descriptor prefetch and the first-operand latency are paid once per
//     layer, and the epilogue of tile i overlaps the main loop of tile i + 1 through the second
//     accumulator.
//
// Warp roles (384 threads): warp 0 activation producer, warp 1 weight producer, warp 2 MMA issuer +
// TMEM owner, warp 3 idle, warps 4-11 epilogue (two per TMEM lane quarter, interleaved 32-column chunks).
#include "conv.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace rmr {

namespace {

constexpr int kThreads2 = 384;
constexpr int kMaxA = 8;     // activation ring slots
constexpr int kMaxB = 40;    // weight slots (resident mode: one per k-block of the CTA)
constexpr int kSmemMax = 224 * 1024;   // dynamic; ~2 KB of static shared memory (barriers, bias) sit beside it
constexpr bool kConv2Default = false;   // until the parity run on the B200 is green

__device__ __forceinline__ void pdl_wait2() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch2() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (spins > (1u << 22)) {   // a protocol bug becomes an error the host sees, never a hang
            printf("rmr conv2: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

__device__ __forceinline__ bool elect_one2() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx2(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kHalo: 3x3 stride-1 layers, one [18][pw] pixel patch per channel chunk serves the nine taps.
template <bool kHalo, bool kRes>
__global__ void __launch_bounds__(kThreads2, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __grid_constant__ Conv2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kMaxA], bar_aempty[kMaxA];
    __shared__ __align__(8) uint64_t bar_bfull[kMaxB], bar_bempty[kMaxB];
    __shared__ __align__(8) uint64_t bar_accfull[2], bar_accempty[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int4 s_tap[9];
    __shared__ float s_bias[256];
    __shared__ int s_last;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // this CTA's slice: output channels [ch0, ch0 + block_n), k-blocks of split `split`, pixel tiles j, j + gm, ...
    const int ns = blockIdx.x % p.ns_total;
    const int j0 = blockIdx.x / p.ns_total;
    const int nt = ns / p.splits;
    const int split = ns - nt * p.splits;
    const int ch0 = nt * p.block_n;
    // k-block range of this split: per-tap mode counts (tap, chunk) pairs tap-major, halo mode counts chunks
    const int u0 = split * p.units_per_split;
    const int u1 = min(u0 + p.units_per_split, kHalo ? p.kpt : p.ntaps * p.kpt);

    if (warp == 1) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
        }
        if (lane < p.ntaps) s_tap[lane] = p.tap[lane];
    } else if (warp == 2) {
        if (lane == 0) {
            for (int i = 0; i < p.sa; ++i) {
                mbar_init(smem_u32(&bar_afull[i]), 1);
                mbar_init(smem_u32(&bar_aempty[i]), 1);
            }
            for (int i = 0; i < p.sb; ++i) {
                mbar_init(smem_u32(&bar_bfull[i]), 1);
                mbar_init(smem_u32(&bar_bempty[i]), 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(smem_u32(&bar_accfull[i]), 1);
                mbar_init(smem_u32(&bar_accempty[i]), 8);   // one arrival per epilogue warp
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
        tmem_relinquish();
    } else if (warp >= 4) {
        const int i = threadIdx.x - 128;
        if (i < p.block_n) s_bias[i] = __ldg(p.bias + ch0 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    pdl_launch2();

    const uint32_t a_ring = smem_base + p.a_off;
    const uint32_t b_ring = smem_base + p.b_off;

    // The three issuing roles run warp-converged: every lane walks the loop and waits on the barriers, one
    // elected lane issues.  Issued from divergent code, ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an
    // ELECT + BRA.U.ANY lane loop (~150 cycles per instruction, tools/umma_bench.cu); with elect.sync under
    // warp-uniform control flow the bookkeeping stays in the uniform datapath and the MMAs issue back to back.
    if (warp == 0) {
        // ------------------------------ activation producer ------------------------------
        const bool leader = elect_one2();
        pdl_wait2();   // activations are the previous layer's output
        // ring bookkeeping without divisions: (slot, phase) counters; a fresh barrier passes a wait on parity 1
        int slot = 0;
        uint32_t phase = 0;
        for (int mt = j0; mt < p.m_tiles; mt += p.gm) {
            int t = mt;
            const int tile_w = t % p.tiles_w; t /= p.tiles_w;
            const int tile_h = t % p.tiles_h;
            const int tile_n = t / p.tiles_h;
            const int ow0 = tile_w * p.tw, oh0 = tile_h * p.th, n0 = tile_n * p.tn;
            if (kHalo) {
                for (int c = u0; c < u1; ++c) {
                    mbar_wait_spin(smem_u32(&bar_aempty[slot]), phase ^ 1u);
                    if (leader) {
                        const uint32_t full = smem_u32(&bar_afull[slot]);
                        mbar_expect_tx(full, p.a_tx);
                        tma_load_4d(a_ring + slot * p.a_stage, &tm_a, full, p.cin_coff + c * p.bk, ow0 - 1, oh0 - 1, n0);
                    }
                    __syncwarp();
                    if (++slot == p.sa) { slot = 0; phase ^= 1u; }
                }
            } else {
                int tap = u0 / p.kpt, kc = u0 - tap * p.kpt;
                for (int u = u0; u < u1; ++u) {
                    mbar_wait_spin(smem_u32(&bar_aempty[slot]), phase ^ 1u);
                    const int4 tp = s_tap[tap];
                    if (leader) {
                        const uint32_t full = smem_u32(&bar_afull[slot]);
                        mbar_expect_tx(full, p.a_tx);
                        tma_load_5d(a_ring + slot * p.a_stage, &tm_a, full, p.cin_coff + tp.x + kc * p.bk, ow0 + tp.y, tp.z,
                                    oh0 + tp.w, n0);
                    }
                    __syncwarp();
                    if (++kc == p.kpt) { kc = 0; ++tap; }
                    if (++slot == p.sa) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ weight producer ------------------------------
        // weights are constants: no grid dependency.  Resident mode loads the CTA's slice once.
        const bool leader = elect_one2();
        int slot = 0;
        uint32_t phase = 0;
        for (int mt = j0; mt < p.m_tiles; mt += p.gm) {
            if (p.b_resident && mt != j0) break;
            if (kHalo) {
                for (int c = u0; c < u1; ++c)
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait_spin(smem_u32(&bar_bempty[slot]), phase ^ 1u);
                        if (leader) {
                            const uint32_t full = smem_u32(&bar_bfull[slot]);
                            mbar_expect_tx(full, p.b_stage);
                            tma_load_2d(b_ring + slot * p.b_stage, &tm_b, full, tap * p.cin + c * p.bk, ch0);
                        }
                        __syncwarp();
                        if (++slot == p.sb) { slot = 0; phase ^= 1u; }
                    }
            } else {
                for (int u = u0; u < u1; ++u) {
                    mbar_wait_spin(smem_u32(&bar_bempty[slot]), phase ^ 1u);
                    if (leader) {
                        const uint32_t full = smem_u32(&bar_bfull[slot]);
                        mbar_expect_tx(full, p.b_stage);
                        tma_load_2d(b_ring + slot * p.b_stage, &tm_b, full, u * p.bk, ch0);
                    }
                    __syncwarp();
                    if (++slot == p.sb) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ------------------------------ MMA issuer ------------------------------
        const bool leader = elect_one2();
        const uint64_t adesc0 = umma_smem_desc(0, p.sbo_a, p.layout);
        const uint64_t bdesc0 = umma_smem_desc(0, p.sbo_b, p.layout);
        const int ksteps = p.bk >> 4;
        uint32_t it = 0;
        int aslot = 0, bslot = 0;
        uint32_t aphase = 0, bphase = 0;
        for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++it) {
            const uint32_t ab = it & 1u;
            mbar_wait_spin(smem_u32(&bar_accempty[ab]), ((it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t acc = tmem_base + ab * p.acc_stride;
            uint32_t accumulate = 0;
            if (kHalo) {
                if (p.b_resident) { bslot = 0; bphase = 0; }   // resident weights: slot = k-block index, loaded once
                for (int c = u0; c < u1; ++c) {
                    mbar_wait_spin(smem_u32(&bar_afull[aslot]), aphase);
                    const uint32_t a_addr = a_ring + aslot * p.a_stage;
                    int dy = 0, dx = 0;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!p.b_resident || it == 0) mbar_wait_spin(smem_u32(&bar_bfull[bslot]), bphase);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t ad = adesc0 | (((a_addr + (dy * p.pw + dx) * p.row_bytes) & 0x3FFFF) >> 4);
                            const uint64_t bd = bdesc0 | (((b_ring + bslot * p.b_stage) & 0x3FFFF) >> 4);
                            if (ksteps == 4) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_f16(acc, ad + 2u * k, bd + 2u * k, p.idesc, k ? 1u : accumulate);
                            } else {
#pragma unroll
                                for (int k = 0; k < 2; ++k) umma_f16(acc, ad + 2u * k, bd + 2u * k, p.idesc, k ? 1u : accumulate);
                            }
                            if (!p.b_resident) umma_commit(smem_u32(&bar_bempty[bslot]));
                            if (tap == 8) umma_commit(smem_u32(&bar_aempty[aslot]));   // patch consumed
                        }
                        __syncwarp();
                        accumulate = 1;
                        if (++dx == 3) { dx = 0; ++dy; }
                        if (++bslot == p.sb) { bslot = 0; bphase ^= 1u; }
                    }
                    if (++aslot == p.sa) { aslot = 0; aphase ^= 1u; }
                }
            } else {
                if (p.b_resident) { bslot = 0; bphase = 0; }
#pragma unroll 1
                for (int u = u0; u < u1; ++u) {
                    if (!p.b_resident || it == 0) mbar_wait_spin(smem_u32(&bar_bfull[bslot]), bphase);
                    mbar_wait_spin(smem_u32(&bar_afull[aslot]), aphase);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t ad = adesc0 | (((a_ring + aslot * p.a_stage) & 0x3FFFF) >> 4);
                        const uint64_t bd = bdesc0 | (((b_ring + bslot * p.b_stage) & 0x3FFFF) >> 4);
                        if (ksteps == 4) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16(acc, ad + 2u * k, bd + 2u * k, p.idesc, k ? 1u : accumulate);
                        } else {
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_f16(acc, ad + 2u * k, bd + 2u * k, p.idesc, k ? 1u : accumulate);
                        }
                        umma_commit(smem_u32(&bar_aempty[aslot]));
                        if (!p.b_resident) umma_commit(smem_u32(&bar_bempty[bslot]));
                    }
                    __syncwarp();
                    accumulate = 1;
                    if (++aslot == p.sa) { aslot = 0; aphase ^= 1u; }
                    if (++bslot == p.sb) { bslot = 0; bphase ^= 1u; }
                }
            }
            if (leader) umma_commit(smem_u32(&bar_accfull[ab]));   // accumulator of this tile complete
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ------------------------------ epilogue: 8 warps, TMEM lane quarter = warp % 4 ------------------------------
        const int q = warp & 3;
        const int chunk0 = warp >= 8 ? 32 : 0;
        const int row = q * 32 + lane;
        const int tw_i = row % p.tw;
        const int th_i = (row / p.tw) % p.th;
        const int tn_i = row / (p.tw * p.th);
        const int nvalid = min(p.block_n, p.cout - ch0);
        const bool vec = p.vec_ok != 0;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        pdl_wait2();   // residual reads and output writes must follow the previous grid
        uint32_t it = 0;
        for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++it) {
            int t = mt;
            const int tile_w = t % p.tiles_w; t /= p.tiles_w;
            const int tile_h = t % p.tiles_h;
            const int tile_n = t / p.tiles_h;
            const int ow = tile_w * p.tw + tw_i, oh = tile_h * p.th + th_i, n = tile_n * p.tn + tn_i;
            const bool valid = (ow < p.w_out) && (oh < p.h_out) && (n < p.n);
            const size_t pix = (static_cast<size_t>(n) * p.h_out + oh) * p.w_out + ow;
            const __half* rptr = (kRes && p.res != nullptr && valid) ? p.res + pix * p.res_pitch + p.res_coff + ch0 : nullptr;
            const uint32_t ab = it & 1u;
            const uint32_t taddr = tmem_base + ab * p.acc_stride + lane_addr;

            uint4 rres[4];
            auto fetch_res = [&](int c0) {
                const int cnt = nvalid - c0;
                if (kRes && rptr != nullptr && vec && cnt >= 16) {
                    rres[0] = *reinterpret_cast<const uint4*>(rptr + c0);
                    rres[1] = *reinterpret_cast<const uint4*>(rptr + c0 + 8);
                    if (cnt >= 32) {
                        rres[2] = *reinterpret_cast<const uint4*>(rptr + c0 + 16);
                        rres[3] = *reinterpret_cast<const uint4*>(rptr + c0 + 24);
                    }
                }
            };
            auto finish_chunk = [&](int c0, float (&f)[32]) {
                const int cnt = min(32, nvalid - c0);
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] += s_bias[c0 + j];
                if (p.act == 1) {
                    // SiLU = x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU per element
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float h = 0.5f * f[j];
                        f[j] = fmaf(h, tanh_approx(h), h);
                    }
                } else if (p.act == 2) {
                    // SiLU through ex2 + rcp (two MUFU per element), staged so 16 chains pipeline
#pragma unroll
                    for (int h0 = 0; h0 < 32; h0 += 16) {
                        float e[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) e[j] = 1.0f + ex2_approx(-1.4426950408889634f * f[h0 + j]);
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[h0 + j] = f[h0 + j] * rcp_approx2(e[j]);
                    }
                }
                if (kRes && rptr != nullptr) {
                    if (vec && cnt >= 16) {
                        const __half2* h = reinterpret_cast<const __half2*>(rres);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 a = __half22float2(h[j]);
                            f[2 * j] += a.x; f[2 * j + 1] += a.y;
                        }
                        if (cnt >= 32) {
#pragma unroll
                            for (int j = 8; j < 16; ++j) {
                                const float2 a = __half22float2(h[j]);
                                f[2 * j] += a.x; f[2 * j + 1] += a.y;
                            }
                        } else {
#pragma unroll
                            for (int j = 16; j < 32; ++j)
                                if (j < cnt) f[j] += __half2float(rptr[c0 + j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < cnt) f[j] += __half2float(rptr[c0 + j]);
                    }
                }
                if (p.out_f32) {
                    float* o = static_cast<float*>(p.out) + pix * p.out_pitch + p.out_coff + ch0 + c0;
                    if (vec && (cnt & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (4 * j < cnt)
                                reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < cnt) o[j] = f[j];
                    }
                } else {
                    __half* o = static_cast<__half*>(p.out) + pix * p.out_pitch + p.out_coff + ch0 + c0;
                    if (vec && (cnt & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (8 * j < cnt) {
                                uint4 w;
                                __half2* h = reinterpret_cast<__half2*>(&w);
#pragma unroll
                                for (int u = 0; u < 4; ++u) h[u] = __floats2half2_rn(f[8 * j + 2 * u], f[8 * j + 2 * u + 1]);
                                reinterpret_cast<uint4*>(o)[j] = w;
                                if (p.dup_mode == 1) {
                                    reinterpret_cast<uint4*>(p.dup + pix * p.dup_pitch + p.dup_coff + ch0 + c0)[j] = w;
                                } else if (p.dup_mode == 2) {
                                    // nearest 2x upsample written by the producer: (oh, ow) -> (2oh + dy, 2ow + dx)
                                    const size_t up = (static_cast<size_t>(n) * (2 * p.h_out) + 2 * oh) * (2 * p.w_out) + 2 * ow;
                                    __half* u0p = p.dup + up * p.dup_pitch + p.dup_coff + ch0 + c0;
                                    __half* u1p = u0p + static_cast<size_t>(2 * p.w_out) * p.dup_pitch;
                                    reinterpret_cast<uint4*>(u0p)[j] = w;
                                    reinterpret_cast<uint4*>(u0p + p.dup_pitch)[j] = w;
                                    reinterpret_cast<uint4*>(u1p)[j] = w;
                                    reinterpret_cast<uint4*>(u1p + p.dup_pitch)[j] = w;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < cnt) o[j] = __float2half_rn(f[j]);
                    }
                }
            };

            mbar_wait_spin(smem_u32(&bar_accfull[ab]), (it >> 1) & 1);
            tc_fence_after();
            // the last chunk this warp reads: once it is in registers the accumulator goes back to the issuer
            int last_c0 = -1;
            for (int c0 = chunk0; c0 < nvalid; c0 += 64) last_c0 = c0;
            if (p.splits == 1) {
                for (int c0 = chunk0; c0 < nvalid; c0 += 64) {
                    uint32_t v[32];
                    __syncwarp();   // tcgen05.ld is warp-aligned: reconverge after the predicated stores
                    tmem_ld_32(taddr + c0, v);
                    fetch_res(c0);
                    tmem_ld_wait();
                    if (c0 == last_c0) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                    }
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (valid) finish_chunk(c0, f);
                }
                if (last_c0 < 0) {   // this warp has no chunk (N <= 32 and warp >= 8): still release the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                }
            } else {
                // ---- split-K: every split parks its raw fp32 partial tile; the split that arrives last sums
                // them in split order (deterministic) and runs the real epilogue ----
                const size_t tile_lin = static_cast<size_t>(nt) * p.m_tiles + mt;
                float* part = p.partial + (tile_lin * p.splits * 128 + row) * p.part_ld;
                const size_t split_stride = static_cast<size_t>(128) * p.part_ld;
                for (int c0 = chunk0; c0 < nvalid; c0 += 64) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
                        float4* o = reinterpret_cast<float4*>(part + split * split_stride + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                asm volatile("bar.sync 1, 256;" ::: "memory");   // the epilogue warps: all partial stores issued
                if (threadIdx.x == 128) {
                    __threadfence();   // cumulative: publishes the CTA's stores ordered before it by the barrier
                    s_last = (atomicAdd(p.counters + tile_lin, 1) == p.splits - 1) ? 1 : 0;
                    __threadfence();
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (s_last) {
                    for (int c0 = chunk0; c0 < nvalid; c0 += 64) {
                        fetch_res(c0);
                        if (valid) {
                            float f[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) f[j] = 0.f;
                            for (int z = 0; z < p.splits; ++z) {
                                const float4* src = reinterpret_cast<const float4*>(part + z * split_stride + c0);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 tv = __ldcg(src + j);
                                    f[4 * j] += tv.x; f[4 * j + 1] += tv.y; f[4 * j + 2] += tv.z; f[4 * j + 3] += tv.w;
                                }
                            }
                            finish_chunk(c0, f);
                        }
                    }
                    if (threadIdx.x == 128) p.counters[tile_lin] = 0;   // ready for the next launch of this layer
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");   // s_last is rewritten by the next tile
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

bool env_flag(const char* name, bool dflt) {
    const char* e = std::getenv(name);
    if (!e || !e[0]) return dflt;
    return e[0] == '1';
}
int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    if (!e || !e[0]) return dflt;
    return std::atoi(e);
}

using EncodeTiledFn2 = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn2 encode_fn2() {
    static EncodeTiledFn2 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        RMR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || ptr == nullptr)
            throw CudaError("cuTensorMapEncodeTiled driver entry point not available");
        fn = reinterpret_cast<EncodeTiledFn2>(ptr);
    });
    return fn;
}
void encode2(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box, CUtensorMapSwizzle swz) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn2()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides_bytes, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
}

// Cost model (cycles) of one layer under a candidate configuration; constants from tools/umma_bench.cu
// (profiles/r2_umma_bench.txt).  Used only to rank configurations.
struct Cand {
    int block_n, splits, halo;
    double cost;
};

}  // namespace

bool conv2_enabled() {
    // RMR_CONV_V2=0 selects the round-1 kernel (conv.cu)
    static const bool on = env_flag("RMR_CONV_V2", kConv2Default);
    return on;
}

bool conv2_supported(const ConvDesc& d) {
    if (!conv_umma_supported(d)) return false;
    return true;
}

// Host-only part: tile shape, channel tile, split-K, shared-memory layout (no CUDA calls: also used by the
// plan dump of tools/conv_plan.py on a machine without a GPU).
void plan_conv2(const ConvDesc& d, ConvLaunch& l) {
    Conv2Params& p = l.q;
    std::memset(&p, 0, sizeof(p));
    l.v2 = 1;
    p.n = d.n; p.h_out = d.h_out; p.w_out = d.w_out; p.cout = d.cout;
    p.bk = (d.cin % 64 == 0) ? 64 : 32;
    p.kpt = d.cin / p.bk;
    p.ntaps = d.k * d.k;
    p.cin = d.cin; p.cin_coff = d.in_coff;
    p.row_bytes = p.bk * 2;
    const int S = d.stride;

    // ---- A mode: halo patch for 3x3 stride-1 layers on maps large enough for 8 x 16 pixel tiles ----
    const int halo_min = env_int("RMR_HALO_MIN", 40);
    const bool halo_ok = d.k == 3 && S == 1 && d.h_out >= halo_min && d.w_out >= halo_min &&
                         (p.bk == 64 || env_flag("RMR_HALO32", true));
    p.halo = (env_flag("RMR_HALO", true) && halo_ok) ? 1 : 0;
    if (p.halo) {
        p.tw = 8; p.th = 16; p.tn = 1;
        p.pw = env_int("RMR_HALO_PW", 10);
    } else {
        long best = -1;
        for (int tw = 128; tw >= 1; tw >>= 1)
            for (int th = 128 / tw; th >= 1; th >>= 1) {
                const int tn = 128 / (tw * th);
                const long tiles = static_cast<long>((d.w_out + tw - 1) / tw) * ((d.h_out + th - 1) / th) * ((d.n + tn - 1) / tn);
                if (best < 0 || tiles < best) { best = tiles; p.tw = tw; p.th = th; p.tn = tn; }
            }
    }
    p.tiles_w = (d.w_out + p.tw - 1) / p.tw;
    p.tiles_h = (d.h_out + p.th - 1) / p.th;
    p.tiles_n = (d.n + p.tn - 1) / p.tn;
    p.m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;

    // ---- output-channel tile and split-K: rank the candidates with the cost model ----
    const int units = p.halo ? p.kpt : p.ntaps * p.kpt;          // split granularity (halo: channel chunks)
    const int kb_per_unit = p.halo ? 9 : 1;
    const int ksteps = p.bk / 16;
    const int force_n = env_int("RMR_CONV_N", 0);
    const int force_splits = env_int("RMR_CONV_SPLITS", 0);
    const int kSM = 148;
    double best_cost = -1;
    int best_n = 0, best_splits = 1;
    for (int bn = 256; bn >= 16; bn -= 16) {
        if (d.cout_pad % bn != 0) continue;
        if (force_n && bn != force_n && d.cout_pad % force_n == 0) continue;
        const int n_tiles = d.cout_pad / bn;
        // split-K (deterministic: fp32 partial tiles through L2, last-arriving split reduces) stays available for
        // experiments (RMR_CONV_SPLITS=n) but is not planned: a partial tile is 128 x N x 4 B written and read
        // once per split — for the layers that would want it that is 10-30x the layer's own traffic (measured:
        // 40x40x256 -> 384 stride 2 at batch 7, 8 splits: 228 us against 25 us for the round-1 kernel).  Narrower
        // channel tiles give the small maps their parallelism instead: below N = 64 an MMA costs the same ~56
        // issue cycles whatever N is, so more, narrower CTAs are free until the activation re-reads show.
        const int max_splits = force_splits ? force_splits : 1;
        for (int splits = 1; splits <= max_splits; ++splits) {
            if (force_splits && splits != force_splits && units >= force_splits) continue;
            const int ups = (units + splits - 1) / splits;
            if ((units + ups - 1) / ups != splits) continue;       // no empty split
            const long ns = static_cast<long>(n_tiles) * splits;
            const long gm = std::max<long>(1, std::min<long>(p.m_tiles, kSM / std::max<long>(1, std::min<long>(ns, kSM))));
            const long ctas = ns * gm;
            const double waves = std::ceil(static_cast<double>(ctas) / kSM);
            const double busy = static_cast<double>(std::min<long>(ctas, kSM));
            const double tiles_per_cta = std::ceil(static_cast<double>(p.m_tiles) / gm);
            const int kb = ups * kb_per_unit;
            // one K = 16 slice (constants measured with tools/umma_bench.cu, profiles/r2_umma_bench.txt):
            //   issue   ~56 clk per tcgen05.mma from one issuing thread (groups of 4 + wait + commit)
            //   tensor  N / 2 clk
            //   smem    (A 4096 B + B 32 N B read by the MMA + bytes TMA writes for the slice) / 128 B/clk
            //   ingest  L2 -> SM: 125 B/clk for a lone SM, ~75 B/clk per SM when all 148 stream
            const double a_write = p.halo ? 4096.0 * (18.0 * p.pw) / (9.0 * 128.0) : 4096.0;
            const bool resident = static_cast<double>(kb) * bn * p.bk * 2 + (p.halo ? 2.0 : 4.0) * 16384 <= kSmemMax - 1024 && kb <= kMaxB;
            const double b_write = resident ? bn * 32.0 / tiles_per_cta : bn * 32.0;
            const double smem_cyc = (4096.0 + bn * 32.0 + a_write + b_write) / 128.0;
            const double ingest_rate = std::min(125.0, 11000.0 / busy);
            const double slice = std::max({56.0, bn / 2.0, smem_cyc, (a_write + b_write) / ingest_rate});
            const double per_kb = ksteps * slice;
            // epilogue of one tile: 8 warps, 32-column chunks (TMEM load, bias, SiLU, fp16 stores); overlaps the next
            // tile's main loop, so a tile costs the larger of the two
            const double epi = 500.0 + (bn > 64 ? (bn - 64) * 7.0 : 0.0) + (splits > 1 ? 4000.0 + bn * 30.0 * splits : 0.0);
            const double tile_cyc = std::max(kb * per_kb, epi);
            // per-CTA fixed cost: launch ramp, barrier init, TMEM allocation, first operands in flight, last epilogue
            const double cost = waves * (2400.0 + tiles_per_cta * tile_cyc + epi);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_n = bn; best_splits = splits; }
        }
    }
    if (best_n == 0) throw CudaError("conv2: no output-channel tile for cout_pad " + std::to_string(d.cout_pad));
    p.block_n = best_n;
    p.n_tiles = d.cout_pad / p.block_n;
    p.splits = best_splits;
    p.units_per_split = (units + p.splits - 1) / p.splits;
    p.ns_total = p.n_tiles * p.splits;
    p.gm = static_cast<int>(std::max<long>(1, std::min<long>(p.m_tiles, kSM / std::max(1, std::min(p.ns_total, kSM)))));
    const int kb_cta = p.units_per_split * kb_per_unit;

    // ---- shared-memory layout: [A ring][B slots] ----
    p.b_stage = static_cast<uint32_t>(p.block_n) * p.bk * 2u;
    if (p.halo) {
        p.a_tx = 18u * p.pw * p.row_bytes;
        p.a_stage = (p.a_tx + 1023u) & ~1023u;
    } else {
        p.a_tx = 128u * p.row_bytes;
        p.a_stage = p.a_tx;
    }
    const int budget = kSmemMax - 1024;
    const int tiles_per_cta = (p.m_tiles + p.gm - 1) / p.gm;
    const int a_units_total = tiles_per_cta * p.units_per_split;          // A stages this CTA ever loads
    const int a_min = std::min(a_units_total, p.halo ? 2 : 4);
    p.b_resident = (kb_cta <= kMaxB && static_cast<long>(kb_cta) * p.b_stage + static_cast<long>(a_min) * p.a_stage <= budget &&
                    env_flag("RMR_B_RESIDENT", true)) ? 1 : 0;
    if (p.b_resident) {
        p.sb = kb_cta;
        p.sa = static_cast<int>(std::min<long>({kMaxA, a_units_total, (budget - static_cast<long>(p.sb) * p.b_stage) / p.a_stage}));
    } else {
        // streaming weights.  Halo mode: two or three patch slots, the rest of the budget is weight slots;
        // per-tap mode: A and B are consumed in lock step, so both rings cover the same number of k-blocks.
        const long total_kb = static_cast<long>(kb_cta) * tiles_per_cta;
        if (p.halo) {
            const long a_res = static_cast<long>(std::min(a_units_total, 3)) * p.a_stage;
            p.sb = static_cast<int>(std::max<long>(2, std::min<long>({kMaxB, (budget - a_res) / p.b_stage, total_kb})));
        } else {
            p.sb = static_cast<int>(std::max<long>(2, std::min<long>({kMaxB, budget / (p.a_stage + p.b_stage), total_kb})));
        }
        p.sa = static_cast<int>(std::max<long>(1, std::min<long>({kMaxA, a_units_total, (budget - static_cast<long>(p.sb) * p.b_stage) / p.a_stage})));
    }
    if (p.sa < 1 || static_cast<long>(p.sa) * p.a_stage + static_cast<long>(p.sb) * p.b_stage > budget)
        throw CudaError("conv2: operand ring does not fit in shared memory");
    p.a_off = 0;
    p.b_off = static_cast<uint32_t>(p.sa) * p.a_stage;
    l.smem_bytes = static_cast<int>(p.b_off + static_cast<uint32_t>(p.sb) * p.b_stage) + 1024;

    p.acc_stride = static_cast<uint32_t>((p.block_n + 31) / 32 * 32);
    const uint32_t need = 2 * p.acc_stride;
    p.tmem_cols = need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
    p.part_ld = static_cast<int>(p.acc_stride);

    const int out_align = d.out_f32 ? 4 : 8;
    p.vec_ok = (d.out_pitch % out_align == 0 && d.out_coff % out_align == 0 &&
                (d.res == nullptr || (d.res_pitch % 8 == 0 && d.res_coff % 8 == 0))) ? 1 : 0;
    for (int r = 0; r < d.k; ++r)
        for (int s = 0; s < d.k; ++s) {
            int4 t;
            if (d.k == 1) t = make_int4(0, 0, 0, 0);
            else if (S == 1) t = make_int4(0, s - 1, 0, r - 1);
            else t = make_int4((s == 1 ? 0 : 1) * d.in_pitch, s == 0 ? -1 : 0, r == 1 ? 0 : 1, r == 0 ? -1 : 0);
            p.tap[r * d.k + s] = t;
        }
    p.out = d.out; p.out_pitch = d.out_pitch; p.out_coff = d.out_coff; p.out_f32 = d.out_f32;
    p.bias = d.bias;
    p.act = d.act ? (env_flag("RMR_SILU_EXP", false) ? 2 : 1) : 0;
    p.res = d.res; p.res_pitch = d.res_pitch; p.res_coff = d.res_coff;
    p.dup = d.dup; p.dup_pitch = d.dup_pitch; p.dup_coff = d.dup_coff; p.dup_mode = d.dup_mode;
    p.idesc = (1u << 4) | (static_cast<uint32_t>(p.block_n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    p.layout = (p.bk == 64) ? 2u : 4u;
    p.sbo_b = 8u * p.row_bytes;
    p.sbo_a = p.halo ? static_cast<uint32_t>(p.pw) * p.row_bytes : 8u * p.row_bytes;
    l.grid = dim3(static_cast<unsigned>(p.ns_total) * p.gm, 1, 1);
    l.flops = 2.0 * d.n * d.h_out * d.w_out * static_cast<double>(d.cout) * d.k * d.k * d.cin;
    // mirror what the generic plan code reads from the r1 parameter block
    l.p.splits = p.splits; l.p.block_n = p.block_n; l.p.part_ld = p.part_ld; l.p.vec_ok = p.vec_ok;
    l.p.halo = p.halo; l.p.slim = 0; l.p.pair = 0;
}

void make_conv2_launch(const ConvDesc& d, ConvLaunch& l) {
    plan_conv2(d, l);
    Conv2Params& p = l.q;
    const int S = d.stride;
    const CUtensorMapSwizzle swz = (p.bk == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

    const cuuint64_t cp = static_cast<cuuint64_t>(d.in_pitch);
    if (p.halo) {
        cuuint64_t dims[4] = {cp, static_cast<cuuint64_t>(d.w_in), static_cast<cuuint64_t>(d.h_in), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[3] = {cp * 2, d.w_in * cp * 2, static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.pw), 18u, 1u};
        encode2(&l.tm_a, const_cast<__half*>(d.in), 4, dims, strides, box, swz);
    } else {
        cuuint64_t dims[5] = {S * cp, static_cast<cuuint64_t>(d.w_in / S), static_cast<cuuint64_t>(S),
                              static_cast<cuuint64_t>(d.h_in / S), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[4] = {S * cp * 2, d.w_in * cp * 2, S * d.w_in * cp * 2, static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        cuuint32_t box[5] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.tw), 1u, static_cast<cuuint32_t>(p.th),
                             static_cast<cuuint32_t>(p.tn)};
        encode2(&l.tm_a, const_cast<__half*>(d.in), 5, dims, strides, box, swz);
    }
    {
        const cuuint64_t ktot = static_cast<cuuint64_t>(p.ntaps) * d.cin_pad;
        cuuint64_t dims[2] = {ktot, static_cast<cuuint64_t>(d.cout_pad)};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.block_n)};
        encode2(&l.tm_b, const_cast<__half*>(d.w), 2, dims, strides, box, swz);
    }
}

size_t conv2_scratch_bytes(const ConvLaunch& l) {
    const Conv2Params& p = l.q;
    if (p.splits <= 1) return 0;
    const size_t tiles = static_cast<size_t>(p.m_tiles) * p.n_tiles;
    return (tiles * sizeof(int) + 255) / 256 * 256 + tiles * p.splits * 128 * p.part_ld * sizeof(float);
}

void conv2_bind_scratch(ConvLaunch& l, void* zeroed_base) {
    Conv2Params& p = l.q;
    if (p.splits <= 1) return;
    const size_t tiles = static_cast<size_t>(p.m_tiles) * p.n_tiles;
    p.counters = static_cast<int*>(zeroed_base);
    p.partial = reinterpret_cast<float*>(static_cast<char*>(zeroed_base) + (tiles * sizeof(int) + 255) / 256 * 256);
}

void conv2_init() {
    // function attributes are per device: one flag per device ordinal
    static std::once_flag once[64];
    int dev = 0;
    RMR_CUDA(cudaGetDevice(&dev));
    std::call_once(once[dev & 63], [] {
        auto prep = [](auto kernel) {
            RMR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
            RMR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        };
        prep(conv2_kernel<false, false>);
        prep(conv2_kernel<false, true>);
        prep(conv2_kernel<true, false>);
        prep(conv2_kernel<true, true>);
        encode_fn2();
    });
}

void launch_conv2(const ConvLaunch& l, cudaStream_t s, bool pdl) {
    conv2_init();
    static const bool use_pdl = !env_flag("RMR_NO_PDL", false);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = l.grid;
    cfg.blockDim = dim3(kThreads2);
    cfg.dynamicSmemBytes = static_cast<size_t>(l.smem_bytes);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (use_pdl && pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const bool res = l.q.res != nullptr;
    if (l.q.halo) {
        if (res) RMR_CUDA((cudaLaunchKernelEx(&cfg, conv2_kernel<true, true>, l.tm_a, l.tm_b, l.q)));
        else RMR_CUDA((cudaLaunchKernelEx(&cfg, conv2_kernel<true, false>, l.tm_a, l.tm_b, l.q)));
    } else {
        if (res) RMR_CUDA((cudaLaunchKernelEx(&cfg, conv2_kernel<false, true>, l.tm_a, l.tm_b, l.q)));
        else RMR_CUDA((cudaLaunchKernelEx(&cfg, conv2_kernel<false, false>, l.tm_a, l.tm_b, l.q)));
    }
}

}  // namespace rmr
