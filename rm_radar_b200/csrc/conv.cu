// Implicit-GEMM convolution on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA) for the YOLOv8 Conv+bias+SiLU(+shortcut) blocks that the reference runs inside
// TensorRT (/root/reference/src/detect/detector.h:122; FP16 engine built at detector.cpp:223-231).
//
// GEMM view:  D[M = N*Ho*Wo pixels, Cout] = A[M, K = taps*Cin] * W[Cout, K]^T
//   * A is never materialised: for every filter tap one 5-D TMA box load fetches a
//     [TN images][TH rows][TW cols][BK channels] patch of the NHWC activation at the tap's spatial
//     offset.  Out-of-bounds box elements are zero-filled by TMA, which *is* the conv padding.
//     Stride-2 layers view the activation as [N][H/2][2][W/2][2*Cpitch] (row/column parity split),
//     so the same plain tiled TMA serves them: the column parity is a channel offset, the row
//     parity is box coordinate 2.
//   * the TMA box lands in shared memory as 128 pixel rows x BK fp16 (128 B or 64 B per row,
//     hardware swizzled), which is exactly the K-major SWIZZLE_128B / SWIZZLE_64B UMMA operand.
//   * W tiles [BLOCK_N][BK] come from a 2-D tensor map over the packed [Cout_pad][taps*Cin] weights.
//   * warp roles, issue streams, kernel variants and the optional modes (halo tile, split-K, CTA pair)
//     are described at the kernel below; the epilogue is
//     tcgen05.ld -> +bias -> SiLU -> +residual -> fp16 (or fp32 for the head logits) -> NHWC store at
//     the destination channel offset (Concat is just this offset), optionally to a second destination.
#include "conv.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace rmr {

namespace {

constexpr int kMaxStages = 8;
constexpr int kThreads = 320;        // wide variant: 4 producer/epilogue + MMA + weight producer + 4 epilogue-only warps
constexpr int kThreadsSlim = 192;    // slim variant: no epilogue-only warps, three CTAs per SM
constexpr int kSmemBudget = 100 * 1024;   // operand ring per CTA: two CTAs co-reside on one SM
constexpr int kSmemBudgetSlim = 72 * 1024;   // slim variant: three CTAs per SM

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// programmatic dependent launch: the next layer's prologue (barrier init, TMEM alloc, bias + weight
// loads) overlaps this layer's tail; activations are only touched after pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float exp2f_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Warp roles: warps 0-3 = activation (A) TMA producers (ring slots round-robin) and, once their loads are
// out, epilogue; warps 6-9 = epilogue only (two warps per TMEM lane quarter, interleaved column chunks);
// warp 4 = MMA issuer (+ TMEM owner; warps 6-8 issue further MMA streams first when the layer runs two
// or four of them); warp 5 = weight (B) TMA producer.  A 5-D UTMALDG costs ~200 issue cycles on the
// issuing thread and a 2-D one ~60 (tools/tma_bench2.cu; issue cost is per warp, not a shared unit),
// while a 128x64x64 MMA block is only 128 tensor cycles: one producer thread cannot feed the tensor core, four can.  Producer and issuer
// loops run warp-converged (one elected lane issues) so the bookkeeping stays in the uniform datapath;
// two CTAs per SM let one CTA's epilogue hide behind the other's main loop.
// kMode 0: one TMA box per filter tap; 1: + split-K; 2: halo tile (3x3 stride 1); 3: CTA pair (cta_group::2).
// kSlim: 192 threads and three CTAs per SM instead of 320 threads and two — for N <= 64 layers with many
// tiles, where what limits an SM is the number of concurrent MMA issue streams, not the epilogue.
template <int kMode, bool kSlim = false, bool kRes = true>   // kRes: the layer has a shortcut operand
__global__ void __launch_bounds__(kSlim ? kThreadsSlim : kThreads, kSlim ? 3 : 2)
conv_umma_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ ConvParams p) {
    constexpr bool kSplit = (kMode == 1);
    constexpr bool kHalo = (kMode == 2);
    constexpr bool kPair = (kMode == 3);
    // CTA pair: two M-tiles (blockIdx.x even / odd) share one MMA stream issued by the even (leader) CTA:
    // one tcgen05.mma covers 256 x N, each CTA stages its own activation tile and half of the weight tile
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const bool is_leader = cta_rank == 0u;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[kMaxStages];
    __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
    __shared__ __align__(8) uint64_t bar_acc;
    __shared__ __align__(8) uint64_t bar_afull[2], bar_aempty[2];   // halo mode: activation patch slots
    __shared__ uint32_t tmem_base_slot;
    __shared__ int4 s_tap[9];
    __shared__ float s_bias[128];
    __shared__ int s_last;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    int mt = blockIdx.x;
    const int tile_w = mt % p.tiles_w;
    mt /= p.tiles_w;
    const int tile_h = mt % p.tiles_h;
    const int tile_n = mt / p.tiles_h;
    const int ow0 = tile_w * p.tw, oh0 = tile_h * p.th, n0 = tile_n * p.tn;
    const int ch0 = blockIdx.y * p.block_n;   // first output channel of this CTA
    long long* dbg = p.dbg ? p.dbg + ((static_cast<size_t>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    // ---- setup: nothing here reads activations, so under PDL it overlaps the previous layer ----
    if (warp == 5) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
        }
        if (lane < p.ntaps) s_tap[lane] = p.tap[lane];
    } else if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < p.stages; ++i) {
                mbar_init(smem_u32(&bar_full[i]), 1);
                mbar_init(smem_u32(&bar_empty[i]), 1);
            }
            mbar_init(smem_u32(&bar_acc), p.dual ? p.dual : 1);
            if (kHalo)
                for (int i = 0; i < 2; ++i) {
                    mbar_init(smem_u32(&bar_afull[i]), 1);
                    mbar_init(smem_u32(&bar_aempty[i]), 1);
                }
            fence_barrier_init();
        }
        __syncwarp();
        if (kPair) {
            tmem_alloc_2sm(smem_u32(&tmem_base_slot), p.tmem_cols);
            tmem_relinquish_2sm();
        } else {
            tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
            tmem_relinquish();
        }
    } else {
        if (static_cast<int>(threadIdx.x) < p.block_n) s_bias[threadIdx.x] = __ldg(p.bias + ch0 + threadIdx.x);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (kPair) cluster_sync_all();   // the peer's barriers exist before anything signals them
    pdl_launch_dependents();
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    // split-K: blockIdx.z owns k-blocks [it0, it0 + num_it) of the taps x channel-chunks sequence
    const int it0 = kSplit ? blockIdx.z * p.it_per_split : 0;
    const int num_it = kSplit ? min(p.it_per_split, p.ntaps * p.kpt - it0) : p.ntaps * p.kpt;

    // One MMA issue stream: k-blocks first, first + step, ... accumulate into TMEM columns [acc, acc + N).
    // With step 2 two warps (4 and 6) each drive half of the k-blocks into their own accumulator and the
    // epilogue adds the two: what bounds a lone CTA is the ~575-cycle issue sequence per k-block
    // (barrier wait + 4 x tcgen05.mma + commit), and two sequences overlap.  The ring has an even number
    // of slots in that case, so a slot always belongs to the same stream (a parity wait cannot tell phases
    // two apart).
    auto mma_stream = [&](int first, int step, uint32_t acc) {
        const bool leader = elect_one();
        const uint64_t adesc0 = umma_smem_desc(smem_base, p.sbo, p.layout);
        const uint64_t bdesc0 = umma_smem_desc(smem_base + p.b_off, p.sbo, p.layout);
        const uint32_t stage_step = p.stage_stride >> 4;   // descriptor start-address units (16 B)
        const int ksteps = p.bk >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < (is_leader ? num_it : 0); ++it) {
            if (it >= first && ((it - first) & (step - 1)) == 0) {
                if (dbg && leader && first == 0 && it < 10) dbg[44 + it] = clock64();
                mbar_wait(smem_u32(&bar_full[stage]), phase);
                tc_fence_after();
                if (leader) {
                    if (dbg && it < 16) dbg[2 + it] = clock64();
                    const uint64_t ad = adesc0 + static_cast<uint64_t>(stage * stage_step);
                    const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * stage_step);
                    // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in the (addr >> 4) field
                    const int kmax = (p.dbg_flags & 1) ? 0 : ((p.dbg_flags & 2) ? 1 : ksteps);
                    for (int k = 0; k < kmax; ++k) {
                        const uint32_t accumulate = (it != first || k != 0) ? 1u : 0u;
                        if (kPair) umma_f16_2sm(tmem_base + acc, ad + 2u * k, bd + 2u * k, p.idesc, accumulate);
                        else umma_f16(tmem_base + acc, ad + 2u * k, bd + 2u * k, p.idesc, accumulate);
                    }
                    // frees the smem slot (in both CTAs of a pair) when the MMAs retire
                    if (kPair) umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);
                    else umma_commit(smem_u32(&bar_empty[stage]));
                }
                __syncwarp();
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader && is_leader) {
            if (kPair) umma_commit_2sm(smem_u32(&bar_acc), 3);   // accumulator complete (both halves of a pair)
            else umma_commit(smem_u32(&bar_acc));
            if (dbg && first == 0) dbg[18] = clock64();
        }
        __syncwarp();
    };

    if (warp == 5) {
        // ------------------------------ B (weight) producer ------------------------------
        // Weights are constants: no grid dependency, the first ring pass goes out immediately.  A tile's
        // complete_tx may land before the A producer's expect_tx (the tx-count is signed).
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        if (!kHalo) {
            for (int it = 0; it < num_it; ++it) {
                if (it >= p.stages) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                if (leader) {
                    if (dbg && it < 10) dbg[54 + it] = clock64();
                    if (kPair)   // this CTA's half of the weight rows; completion is counted on the leader's barrier
                        tma_load_2d_2sm(smem_base + stage * p.stage_stride + p.b_off, &tm_b, smem_u32(&bar_full[stage]),
                                        (it0 + it) * p.bk, ch0 + static_cast<int>(cta_rank) * (p.block_n >> 1));
                    else
                        tma_load_2d(smem_base + stage * p.stage_stride + p.b_off, &tm_b, smem_u32(&bar_full[stage]),
                                    (it0 + it) * p.bk, ch0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        } else {
            // halo mode: the weight ring has its own barriers; k-blocks run channel-chunk major, tap minor,
            // while the packed weights are [tap][cin]
            const uint32_t b_ring = smem_base + p.na * p.a_bytes;
            int it = 0;
            for (int kc = 0; kc < p.kpt; ++kc)
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    if (it >= p.stages) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                    if (leader) {
                        const uint32_t full = smem_u32(&bar_full[stage]);
                        mbar_expect_tx(full, p.b_bytes);
                        tma_load_2d(b_ring + stage * p.b_bytes, &tm_b, full, tap * p.cin + kc * 64, ch0);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
        }
    } else if (warp == 4) {
        // ------------------------------ MMA issuer ------------------------------
        const bool leader = elect_one();
        if (!kHalo) {
            mma_stream(0, p.dual ? p.dual : 1, 0u);
        } else {
            // Halo mode.  The activation patch of one 64-channel chunk is [18 rows][16 cols] pixels x 128 B
            // (SWIZZLE_128B as written by TMA, slot base 1024-aligned).  The A operand of tap (dy, dx) is the
            // same patch read from pixel row dy*16 + dx on: 16 groups of 8 consecutive pixels, one patch row
            // (2048 B) apart = the UMMA stride-byte-offset.  The start is only 128 B aligned; measured on B200
            // (tests/test_gpu_conv.py) the hardware derives the swizzle phase from the absolute shared-memory
            // address bits [7,10), so the descriptor's base_offset field stays 0 (setting it to dx is wrong).
            const uint32_t b_ring = smem_base + p.na * p.a_bytes;
            const uint64_t bdesc0 = umma_smem_desc(b_ring, 1024u, 2u);
            const uint32_t b_step = p.b_bytes >> 4;
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int kc = 0; kc < p.kpt; ++kc) {
                const int slot = kc & (p.na - 1);
                mbar_wait(smem_u32(&bar_afull[slot]), (kc / p.na) & 1);
                const uint32_t a_slot = smem_base + slot * p.a_bytes;
                for (int tap = 0; tap < 9; ++tap, ++it) {
                    mbar_wait(smem_u32(&bar_full[stage]), phase);
                    tc_fence_after();
                    if (leader) {
                        if (dbg && it < 16) dbg[2 + it] = clock64();
                        const int dy = tap / 3, dx = tap - dy * 3;
                        const uint64_t ad = umma_smem_desc(a_slot + (dy * 16 + dx) * 128, 2048u, 2u);
                        const uint64_t bd = bdesc0 + static_cast<uint64_t>(stage * b_step);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(tmem_base, ad + 2u * k, bd + 2u * k, p.idesc, (it | k) != 0);
                        umma_commit(smem_u32(&bar_empty[stage]));
                        if (tap == 8) umma_commit(smem_u32(&bar_aempty[slot]));   // patch consumed
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (kHalo && leader) {
            umma_commit(smem_u32(&bar_acc));   // accumulator complete
            if (dbg) dbg[18] = clock64();
        }
        __syncwarp();
    } else {
        // ---------------- warps 0-3: A (activation) producers ----------------
        // Ring slot s belongs to warp s % 4 for the whole kernel: a slot's owner meets its empty barrier
        // round after round, so it is never more than one phase ahead of it (a parity wait cannot tell
        // phases two apart).
        pdl_wait();   // activations, residual reads and output writes must follow the previous grid
        if (kHalo) {
            if (warp == 0) {
                const bool leader = elect_one();
                for (int kc = 0; kc < p.kpt; ++kc) {
                    const int slot = kc & (p.na - 1);
                    if (kc >= p.na) mbar_wait(smem_u32(&bar_aempty[slot]), ((kc / p.na) & 1) ^ 1u);
                    if (leader) {
                        const uint32_t full = smem_u32(&bar_afull[slot]);
                        mbar_expect_tx(full, p.a_bytes);
                        if (dbg && kc < 16) dbg[24 + kc] = clock64();
                        tma_load_4d(smem_base + slot * p.a_bytes, &tm_a, full, p.cin_coff + kc * 64, ow0 - 1, oh0 - 1, n0);
                    }
                    __syncwarp();
                }
            }
        } else if (warp < 4) {
            const bool leader = elect_one();
            const uint32_t tx = p.a_bytes + p.b_bytes;
            int stage = 0, tap = it0 / p.kpt, kc = it0 - tap * p.kpt;
            uint32_t phase = 0;
            for (int it = 0; it < num_it; ++it) {
                if ((stage & 3) == warp) {
                    const int4 t = s_tap[tap];
                    const uint32_t full = smem_u32(&bar_full[stage]);
                    if (it >= p.stages) mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
                    if (leader) {
                        // a pair's bytes (two activation tiles + two weight halves) are all counted on the
                        // leader's barrier, which only the leader arms
                        if (is_leader) mbar_expect_tx(full, kPair ? 2u * tx : tx);
                        if (dbg && it < 16) dbg[24 + it] = clock64();
                        if (kPair)
                            tma_load_5d_2sm(smem_base + stage * p.stage_stride, &tm_a, full,
                                            p.cin_coff + t.x + kc * p.bk, ow0 + t.y, t.z, oh0 + t.w, n0);
                        else
                            tma_load_5d(smem_base + stage * p.stage_stride, &tm_a, full,
                                        p.cin_coff + t.x + kc * p.bk, ow0 + t.y, t.z, oh0 + t.w, n0);
                    }
                    __syncwarp();
                }
                if (++kc == p.kpt) { kc = 0; ++tap; }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
        // further issue streams: warp 6 (and 7, 8 with four streams), each with its own accumulator
        if (!kSlim && !kHalo && warp >= 6 && warp - 5 < p.dual) mma_stream(warp - 5, p.dual, (warp - 5) * p.acc_stride);
        // ---------------- epilogue: warps 0-3 and 6-9; TMEM lane quarter = warp % 4 ----------------
        // the two warps of a quarter interleave 32-column chunks (0, 64, .. / 32, 96, ..)
        const int q = warp & 3;
        const int chunk0 = warp >= 6 ? 32 : 0;
        constexpr int kChunkStep = kSlim ? 32 : 64;   // slim: one warp per lane quarter takes every chunk
        const int row = q * 32 + lane;
        const int tw_i = row % p.tw;
        const int th_i = (row / p.tw) % p.th;
        const int tn_i = row / (p.tw * p.th);
        const int ow = ow0 + tw_i, oh = oh0 + th_i, n = n0 + tn_i;
        const bool valid = (ow < p.w_out) && (oh < p.h_out) && (n < p.n);
        const size_t pix = (static_cast<size_t>(n) * p.h_out + oh) * p.w_out + ow;
        const int nvalid = min(p.block_n, p.cout - ch0);   // real output channels in this tile
        const __half* rptr = (kRes && p.res != nullptr && valid) ? p.res + pix * p.res_pitch + p.res_coff + ch0 : nullptr;
        const bool vec = p.vec_ok != 0;

        uint4 rnext[4];
        auto fetch_res = [&](int c0) {
            const int cnt = nvalid - c0;
            if (kRes && rptr != nullptr && vec && cnt >= 16) {
                rnext[0] = *reinterpret_cast<const uint4*>(rptr + c0);
                rnext[1] = *reinterpret_cast<const uint4*>(rptr + c0 + 8);
                if (cnt >= 32) {
                    rnext[2] = *reinterpret_cast<const uint4*>(rptr + c0 + 16);
                    rnext[3] = *reinterpret_cast<const uint4*>(rptr + c0 + 24);
                }
            }
        };
        // bias + SiLU + residual + store of one 32-column chunk held in f[]
        auto finish_chunk = [&](int c0, float (&f)[32], const uint4 (&rcur)[4]) {
            const int cnt = min(32, nvalid - c0);
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += s_bias[c0 + j];
            if (p.act) {
                // SiLU, staged so that 16 independent MUFU chains pipeline (ex2 pass, then rcp pass) without
                // doubling the live registers of the whole chunk
#pragma unroll
                for (int h0 = 0; h0 < 32; h0 += 16) {
                    float e[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) e[j] = 1.0f + exp2f_approx(-1.4426950408889634f * f[h0 + j]);
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[h0 + j] = f[h0 + j] * rcp_approx(e[j]);
                }
            }
            if (kRes && rptr != nullptr) {
                if (vec && cnt >= 16) {
                    const __half2* h = reinterpret_cast<const __half2*>(rcur);
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
                                __half* u0 = p.dup + up * p.dup_pitch + p.dup_coff + ch0 + c0;
                                __half* u1 = u0 + static_cast<size_t>(2 * p.w_out) * p.dup_pitch;
                                reinterpret_cast<uint4*>(u0)[j] = w;
                                reinterpret_cast<uint4*>(u0 + p.dup_pitch)[j] = w;
                                reinterpret_cast<uint4*>(u1)[j] = w;
                                reinterpret_cast<uint4*>(u1 + p.dup_pitch)[j] = w;
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

        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (!kSplit) {
            mbar_wait(smem_u32(&bar_acc), 0);
            if (dbg && threadIdx.x == 0) dbg[19] = clock64();
            tc_fence_after();
            for (int c0 = chunk0; c0 < nvalid; c0 += kChunkStep) {
                uint32_t v[32];
                __syncwarp();   // tcgen05.ld is warp-aligned: reconverge after the predicated stores
                tmem_ld_32(taddr + c0, v);
                fetch_res(c0);   // shortcut operand of this chunk: in flight while the accumulator is read
                tmem_ld_wait();
                if (dbg && threadIdx.x == 0 && c0 == chunk0) dbg[40] = clock64();
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                for (int st = 1; st < p.dual; ++st) {   // the other issue streams' accumulators
                    tmem_ld_32(taddr + st * p.acc_stride + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
                }
                if (valid) finish_chunk(c0, f, rnext);
                if (dbg && threadIdx.x == 0 && c0 == chunk0) dbg[42] = clock64();
            }
        } else {
            // ---- split-K: every split parks its raw fp32 partial tile; the split that arrives last
            // sums them in split order (deterministic) and runs the real epilogue ----
            const size_t tile_lin = static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x;
            float* part = p.partial + (tile_lin * p.splits * 128 + row) * p.part_ld;
            const size_t split_stride = static_cast<size_t>(128) * p.part_ld;
            mbar_wait(smem_u32(&bar_acc), 0);
            if (dbg && threadIdx.x == 0) dbg[19] = clock64();
            tc_fence_after();
            for (int c0 = chunk0; c0 < nvalid; c0 += kChunkStep) {
                uint32_t v[32];
                __syncwarp();
                tmem_ld_32(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
                    float4* o = reinterpret_cast<float4*>(part + blockIdx.z * split_stride + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                           __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            }
            if (dbg && threadIdx.x == 0) dbg[43] = clock64();
            asm volatile("bar.sync 1, %0;" ::"n"(kSlim ? 128 : 256) : "memory");   // the epilogue warps: all partial stores issued
            if (threadIdx.x == 0) {
                __threadfence();   // cumulative: publishes the CTA's stores ordered before it by the barrier
                s_last = (atomicAdd(p.counters + tile_lin, 1) == p.splits - 1) ? 1 : 0;
                __threadfence();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kSlim ? 128 : 256) : "memory");
            if (dbg && threadIdx.x == 0) dbg[44] = clock64();
            if (s_last) {
                for (int c0 = chunk0; c0 < nvalid; c0 += kChunkStep) {
                    fetch_res(c0);
                    if (valid) {
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = 0.f;
                        for (int z = 0; z < p.splits; ++z) {
                            const float4* src = reinterpret_cast<const float4*>(part + z * split_stride + c0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t = __ldcg(src + j);
                                f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
                            }
                        }
                        finish_chunk(c0, f, rnext);
                    }
                }
                if (threadIdx.x == 0) p.counters[tile_lin] = 0;   // ready for the next launch of this layer
            }
        }
        __syncwarp();
        if (dbg && threadIdx.x == 0) dbg[20] = clock64();
    }

    tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[21] = clock64();
    if (kPair) cluster_sync_all();   // both CTAs are done with each other's barriers and tensor memory
    if (warp == 4) {
        tc_fence_after();
        if (kPair) tmem_dealloc_2sm(tmem_base, p.tmem_cols);
        else tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------
// Direct convolution on CUDA cores.  Used for the stem (Cin=3, padded to 4: K=36 is too thin for a
// 128-wide MMA tile) and as the on-device checker the tests compare the tcgen05 path against.
// One thread = one output pixel x 8 consecutive output channels.
// ------------------------------------------------------------------------------------------
__global__ void conv_simt_kernel(ConvDesc d) {
    const int groups = (d.cout + 7) / 8;
    const long total = static_cast<long>(d.n) * d.h_out * d.w_out * groups;
    const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = static_cast<int>(idx % groups);
    long pix = idx / groups;
    const int ow = static_cast<int>(pix % d.w_out);
    const int oh = static_cast<int>((pix / d.w_out) % d.h_out);
    const int n = static_cast<int>(pix / (static_cast<long>(d.w_out) * d.h_out));
    const int pad = d.k / 2;
    const int ktot = d.k * d.k * d.cin_pad;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < d.k; ++r) {
        const int ih = oh * d.stride - pad + r;
        if (ih < 0 || ih >= d.h_in) continue;
        for (int s = 0; s < d.k; ++s) {
            const int iw = ow * d.stride - pad + s;
            if (iw < 0 || iw >= d.w_in) continue;
            const __half* ip = d.in + ((static_cast<size_t>(n) * d.h_in + ih) * d.w_in + iw) * d.in_pitch + d.in_coff;
            const __half* wp = d.w + static_cast<size_t>(g * 8) * ktot + (r * d.k + s) * d.cin_pad;
            for (int c = 0; c < d.cin; ++c) {
                const float x = __half2float(ip[c]);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, __half2float(wp[static_cast<size_t>(j) * ktot + c]), acc[j]);
            }
        }
    }
    pix = (static_cast<long>(n) * d.h_out + oh) * d.w_out + ow;
    for (int j = 0; j < 8; ++j) {
        const int co = g * 8 + j;
        if (co >= d.cout) break;
        float x = acc[j] + d.bias[co];
        if (d.act) x = x / (1.0f + __expf(-x));
        if (d.res) x += __half2float(d.res[pix * d.res_pitch + d.res_coff + co]);
        if (d.out_f32) static_cast<float*>(d.out)[pix * d.out_pitch + d.out_coff + co] = x;
        else static_cast<__half*>(d.out)[pix * d.out_pitch + d.out_coff + co] = __float2half_rn(x);
    }
}

// Stem: 3x3 stride-2, Cin = 3 (stored as 4).  One thread = two horizontally adjacent output pixels x
// all COUT channels: the 3x5 input patch lives in registers, each (cout, tap) weight triple is one
// broadcast LDS.128 feeding 6 FMAs.
template <int COUT>
__global__ void __launch_bounds__(128) conv_stem_kernel(ConvDesc d) {
    __shared__ float4 ws[COUT * 9];
    __shared__ float bs[COUT];
    for (int i = threadIdx.x; i < COUT * 9; i += blockDim.x) {
        const uint2 raw = *reinterpret_cast<const uint2*>(d.w + static_cast<size_t>(i) * 4);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        ws[i] = make_float4(a.x, a.y, b.x, 0.f);
    }
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = d.bias[i];
    __syncthreads();
    const int wpairs = (d.w_out + 1) / 2;
    const long total = static_cast<long>(d.n) * d.h_out * wpairs;
    const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = static_cast<int>(idx % wpairs) * 2;
    const int oy = static_cast<int>((idx / wpairs) % d.h_out);
    const int n = static_cast<int>(idx / (static_cast<long>(wpairs) * d.h_out));
    float x[3][5][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int iy = oy * 2 - 1 + r;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int ix = ox * 2 - 1 + c;
            uint2 raw = make_uint2(0u, 0u);
            if (iy >= 0 && iy < d.h_in && ix >= 0 && ix < d.w_in)
                raw = __ldg(reinterpret_cast<const uint2*>(d.in + ((static_cast<size_t>(n) * d.h_in + iy) * d.w_in + ix) * 4));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            x[r][c][0] = a.x; x[r][c][1] = a.y; x[r][c][2] = b.x;
        }
    }
    const bool second = ox + 1 < d.w_out;
    const size_t pix = (static_cast<size_t>(n) * d.h_out + oy) * d.w_out + ox;
    __half* o0 = static_cast<__half*>(d.out) + pix * d.out_pitch + d.out_coff;
    __half* o1 = o0 + d.out_pitch;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 8) {
        float y0[8], y1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a0 = bs[c0 + j], a1 = a0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float4 w = ws[(c0 + j) * 9 + r * 3 + t];
                    a0 = fmaf(x[r][t][0], w.x, a0); a0 = fmaf(x[r][t][1], w.y, a0); a0 = fmaf(x[r][t][2], w.z, a0);
                    a1 = fmaf(x[r][t + 2][0], w.x, a1); a1 = fmaf(x[r][t + 2][1], w.y, a1); a1 = fmaf(x[r][t + 2][2], w.z, a1);
                }
            y0[j] = a0; y1[j] = a1;
        }
        uint4 p0, p1;
        __half2* h0 = reinterpret_cast<__half2*>(&p0);
        __half2* h1 = reinterpret_cast<__half2*>(&p1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h0[j] = __floats2half2_rn(silu(y0[2 * j]), silu(y0[2 * j + 1]));
            h1[j] = __floats2half2_rn(silu(y1[2 * j]), silu(y1[2 * j + 1]));
        }
        *reinterpret_cast<uint4*>(o0 + c0) = p0;
        if (second) *reinterpret_cast<uint4*>(o1 + c0) = p1;
    }
}


// Stem, second version: the same 3x3 stride-2 Cin = 3 (stored as 4) convolution as an im2col GEMM on mma.sync
// (m16n8k16, fp16 operands, fp32 accumulate).  The layer is HBM-bound by nature (armor, batch 7: 23 MB in, 46 MB out,
// 1.2 GFLOP), so there is nothing for tcgen05 to win; what the SIMT version above spends is FP32 issue slots (1728 FMA
// per thread) and broadcast LDS.  One warp = 16 consecutive output pixels of a row x all 32 channels: K = 9 taps x 4
// channels = 36, padded to three k16 steps; an A fragment element pair is one 32-bit load of a pixel's channel pair,
// the weights sit in 24 registers per thread for the whole kernel, the tile leaves through a padded shared-memory
// transpose as 16-byte stores (64 contiguous bytes per pixel).
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) conv_stem_mma_kernel(ConvDesc d, int tiles_w, long total_tiles) {
    __shared__ __align__(16) unsigned char s_tile[4][16 * 80];   // per warp: 16 pixels x (32 ch fp16 + 16 B pad)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q = lane & 3;
    // weights: B[k][n] = w[n][k], k = tap * 4 + ch; fragment registers of kstep s, n-tile j
    uint32_t bw[3][4][2];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half* wr = d.w + static_cast<size_t>(8 * j + g) * 36;
            const int k0 = 16 * s + 2 * q, k1 = k0 + 8;
            bw[s][j][0] = k0 < 36 ? *reinterpret_cast<const uint32_t*>(wr + k0) : 0u;
            bw[s][j][1] = k1 < 36 ? *reinterpret_cast<const uint32_t*>(wr + k1) : 0u;
        }
    float bias[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bias[j][0] = d.bias[8 * j + 2 * q]; bias[j][1] = d.bias[8 * j + 2 * q + 1]; }
    unsigned char* tile = s_tile[warp];
    const long warps_total = static_cast<long>(gridDim.x) * 4;
    for (long t = static_cast<long>(blockIdx.x) * 4 + warp; t < total_tiles; t += warps_total) {
        const int tx = static_cast<int>(t % tiles_w);
        const long row_id = t / tiles_w;
        const int oy = static_cast<int>(row_id % d.h_out), n = static_cast<int>(row_id / d.h_out);
        const int ox0 = tx * 16;
        const __half* img = d.in + static_cast<size_t>(n) * d.h_in * d.w_in * 4;
        float c[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0.f; }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            uint32_t a[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {            // k half: columns 2q (+8)
                const int k = 16 * s + 2 * q + 8 * h;
                const int tap = k >> 2, ch = k & 3;
                const int iy = 2 * oy - 1 + tap / 3, dxo = tap % 3 - 1;
#pragma unroll
                for (int r = 0; r < 2; ++r) {        // fragment rows g, g + 8
                    const int ix = 2 * (ox0 + g + 8 * r) + dxo;
                    uint32_t v = 0u;
                    if (k < 36 && iy >= 0 && iy < d.h_in && ix >= 0 && ix < d.w_in)
                        v = __ldg(reinterpret_cast<const uint32_t*>(img + (static_cast<size_t>(iy) * d.w_in + ix) * 4 + ch));
                    a[2 * h + r] = v;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_16816(c[j], a, bw[s][j][0], bw[s][j][1]);
        }
        // bias + SiLU (h + h tanh(h), h = x / 2) -> fp16 -> padded shared tile [16][80 B]
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float x0 = c[j][2 * r] + bias[j][0], x1 = c[j][2 * r + 1] + bias[j][1];
                if (d.act) {
                    float h0 = 0.5f * x0, h1 = 0.5f * x1, t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                    x0 = fmaf(h0, t0, h0); x1 = fmaf(h1, t1, h1);
                }
                *reinterpret_cast<__half2*>(tile + (g + 8 * r) * 80 + (8 * j + 2 * q) * 2) = __floats2half2_rn(x0, x1);
            }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int prow = g + 8 * r, ox = ox0 + prow;
            if (ox < d.w_out) {
                const uint4 v = *reinterpret_cast<const uint4*>(tile + prow * 80 + q * 16);
                const size_t pix = (static_cast<size_t>(n) * d.h_out + oy) * d.w_out + ox;
                *reinterpret_cast<uint4*>(static_cast<__half*>(d.out) + pix * d.out_pitch + d.out_coff + q * 8) = v;
            }
        }
        __syncwarp();
    }
}

// ---- memory-bound helpers: 8 channels (16 B) per thread, NHWC ----
__global__ void maxpool5_kernel(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch,
                                int out_coff, int n, int h, int w, int c8) {
    const long total = static_cast<long>(n) * h * w * c8;
    const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = static_cast<int>(idx % c8);
    long pix = idx / c8;
    const int x = static_cast<int>(pix % w);
    const int y = static_cast<int>((pix / w) % h);
    const int b = static_cast<int>(pix / (static_cast<long>(w) * h));
    __half2 m[4];
    const __half2 ninf = __floats2half2_rn(-65504.f, -65504.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ninf;
    for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= h) continue;
        for (int dx = -2; dx <= 2; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= w) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(
                in + ((static_cast<size_t>(b) * h + yy) * w + xx) * in_pitch + in_coff + g * 8);
            const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], hv[j]);
        }
    }
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ho[j] = m[j];
    *reinterpret_cast<uint4*>(out + pix * out_pitch + out_coff + g * 8) = o;
}

// SPPF: y1 = maxpool5(x), y2 = maxpool5(y1), y3 = maxpool5(y2) in one launch.  One block per
// (image, 8-channel group): the H x W x 8 slab stays in shared memory, each pool is a separable
// 5-max (row pass, column pass); the three results go to their channel offsets of the concat buffer.
__global__ void __launch_bounds__(256) sppf_pool3_kernel(const __half* in, int in_pitch, int in_coff, __half* out,
                                                         int out_pitch, int coff1, int coff2, int coff3, int h, int w) {
    extern __shared__ uint4 sp_smem[];
    uint4* cur = sp_smem;
    uint4* tmp = sp_smem + h * w;
    const int g = blockIdx.x, b = blockIdx.y, hw = h * w;
    const size_t base = static_cast<size_t>(b) * hw;
    for (int i = threadIdx.x; i < hw; i += blockDim.x)
        cur[i] = *reinterpret_cast<const uint4*>(in + (base + i) * in_pitch + in_coff + g * 8);
    __syncthreads();
    const int coffs[3] = {coff1, coff2, coff3};
    for (int round = 0; round < 3; ++round) {
        for (int i = threadIdx.x; i < hw; i += blockDim.x) {
            const int x = i % w, y = i / w;
            uint4 m = cur[i];
            __half2* hm = reinterpret_cast<__half2*>(&m);
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (dx == 0 || xx < 0 || xx >= w) continue;
                const uint4 v = cur[y * w + xx];
                const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) hm[j] = __hmax2(hm[j], hv[j]);
            }
            tmp[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < hw; i += blockDim.x) {
            const int x = i % w, y = i / w;
            uint4 m = tmp[i];
            __half2* hm = reinterpret_cast<__half2*>(&m);
#pragma unroll
            for (int dy = -2; dy <= 2; ++dy) {
                const int yy = y + dy;
                if (dy == 0 || yy < 0 || yy >= h) continue;
                const uint4 v = tmp[yy * w + x];
                const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) hm[j] = __hmax2(hm[j], hv[j]);
            }
            *reinterpret_cast<uint4*>(out + (base + i) * out_pitch + coffs[round] + g * 8) = m;
            cur[i] = m;   // each thread rewrites only its own pixels; readers of `cur` wait at the barrier below
        }
        __syncthreads();
    }
}

__global__ void upsample2_kernel(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch,
                                 int out_coff, int n, int h_in, int w_in, int c8) {
    const int h = h_in * 2, w = w_in * 2;
    const long total = static_cast<long>(n) * h * w * c8;
    const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = static_cast<int>(idx % c8);
    const long pix = idx / c8;
    const int x = static_cast<int>(pix % w);
    const int y = static_cast<int>((pix / w) % h);
    const int b = static_cast<int>(pix / (static_cast<long>(w) * h));
    const uint4 v = *reinterpret_cast<const uint4*>(
        in + ((static_cast<size_t>(b) * h_in + (y >> 1)) * w_in + (x >> 1)) * in_pitch + in_coff + g * 8);
    *reinterpret_cast<uint4*>(out + pix * out_pitch + out_coff + g * 8) = v;
}

__global__ void copy_channels_kernel(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch,
                                     int out_coff, long npix, int c8) {
    const long total = npix * c8;
    const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int g = static_cast<int>(idx % c8);
    const long pix = idx / c8;
    *reinterpret_cast<uint4*>(out + pix * out_pitch + out_coff + g * 8) =
        *reinterpret_cast<const uint4*>(in + pix * in_pitch + in_coff + g * 8);
}

// ---- tensor map encoding through the driver entry point (no link-time libcuda dependency) ----
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        RMR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || p == nullptr)
            throw CudaError("cuTensorMapEncodeTiled driver entry point not available");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

void encode(CUtensorMap* tm, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
            const cuuint32_t* box, CUtensorMapSwizzle swz) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = get_encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, base, dims, strides_bytes, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
}

// Split-K is implemented (deterministic: fp32 partial tiles, last-arriving split reduces in split
// order) but OFF by default: on B200 it halves the per-CTA lifetime of the 20x20 / 10x10 layers yet
// leaves the graph replay time unchanged (car 0.445 ms with, 0.436 ms without; profiles/r1_summary.md),
// because those layers already overlap with other graph branches.  RMR_SPLITK=1 turns it on.
bool split_k_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_SPLITK");
        return e && e[0] == '1';
    }();
    return on;
}

// Halo mode is parity-green (tests/test_gpu_conv.py runs it) but OFF by default: it cuts the activation
// bytes of 3x3 layers 4x, yet the main loop is bound by the single-thread MMA issue sequence
// (~135 cycles barrier wait + ~240 commit/loop + ~85 per UTCHMMA, measured with RMR_DBG_FLAGS), not by
// operand bytes, so the replay time does not move (car 0.448 vs 0.435 ms).  RMR_HALO=1 turns it on.
// CTA pairs (cta_group::2, cluster of two M-tiles, one MMA stream for 256 x N) are parity-green but OFF by
// default: measured on B200 a pair's k-block period is ~600 cycles for two tiles, the same per-SM rate two
// independent CTAs on one SM already reach (2 x ~575 concurrently), and the cluster syncs add ~1300 cycles
// of setup — graph replay car 0.529 vs 0.477 ms, armor(7) 1.110 vs 0.982 ms.  RMR_PAIR=1 turns it on.
bool pair_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_PAIR");
        return e && e[0] == '1';
    }();
    return on;
}

bool dual_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_NO_DUAL");
        return !(e && e[0] == '1');
    }();
    return on;
}

bool quad_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_NO_QUAD");
        return !(e && e[0] == '1');
    }();
    return on;
}

bool slim_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_NO_SLIM");
        return !(e && e[0] == '1');
    }();
    return on;
}

bool halo_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("RMR_HALO");
        return e && e[0] == '1';
    }();
    return on;
}

// Output-channel tile: the widest divisor of cout_pad (multiple of 16, <= 128) that still yields at
// least one CTA per SM; small feature maps take narrower tiles (more CTAs, shorter epilogues).
int pick_block_n(int cout_pad, long m_tiles) {
    int best = 0;
    for (int bn = 128; bn >= 16; bn -= 16) {
        if (cout_pad % bn != 0) continue;
        if (best == 0) best = bn;
        if (bn < 32 && best >= 32) break;
        best = bn;
        if (m_tiles * (cout_pad / bn) >= 148) break;
    }
    return best;
}

}  // namespace

bool conv_umma_supported(const ConvDesc& d) {
    if (d.cin % 32 != 0 || d.cin != d.cin_pad) return false;
    if (!((d.k == 1 && d.stride == 1) || (d.k == 3 && (d.stride == 1 || d.stride == 2)))) return false;
    if (d.stride == 2 && ((d.h_in | d.w_in) & 1)) return false;
    if (d.in_pitch % 8 || d.in_coff % 8 || d.cout_pad % 16) return false;
    return true;
}

ConvLaunch make_conv_launch(const ConvDesc& d) {
    if (!conv_umma_supported(d)) throw CudaError("conv shape not supported by the tcgen05 path");
    ConvLaunch l;
    std::memset(&l, 0, sizeof(l));
    if (conv2_enabled() && conv2_supported(d)) {
        make_conv2_launch(d, l);
        return l;
    }
    ConvParams& p = l.p;
    p.n = d.n; p.h_out = d.h_out; p.w_out = d.w_out; p.cout = d.cout;
    p.bk = (d.cin % 64 == 0) ? 64 : 32;
    p.kpt = d.cin / p.bk;
    p.ntaps = d.k * d.k;
    p.cin = d.cin; p.cin_coff = d.in_coff;
    // halo mode: 3x3 stride-1 layers with 64-channel chunks on maps of at least 16x16 read one
    // [18][16]-pixel patch per chunk instead of nine shifted 128-pixel boxes (4x fewer activation bytes
    // through the SM's TMA port, which is what bounds the main loop)
    p.halo = (halo_enabled() && d.k == 3 && d.stride == 1 && d.cin % 64 == 0 && d.h_out >= 16 && d.w_out >= 16) ? 1 : 0;
    if (p.halo) {
        p.tw = 8; p.th = 16; p.tn = 1;
    } else {
    // tile shape: minimise the number of 128-pixel tiles (zero-filled lanes are wasted MMA rows)
    long best = -1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
        for (int th = 128 / tw; th >= 1; th >>= 1) {
            const int tn = 128 / (tw * th);
            const long tiles = static_cast<long>((d.w_out + tw - 1) / tw) * ((d.h_out + th - 1) / th) *
                               ((d.n + tn - 1) / tn);
            if (best < 0 || tiles < best) {
                best = tiles; p.tw = tw; p.th = th; p.tn = tn;
            }
        }
    }
    }
    p.tiles_w = (d.w_out + p.tw - 1) / p.tw;
    p.tiles_h = (d.h_out + p.th - 1) / p.th;
    p.tiles_n = (d.n + p.tn - 1) / p.tn;
    p.block_n = pick_block_n(d.cout_pad, static_cast<long>(p.tiles_w) * p.tiles_h * p.tiles_n);
    // operand ring: A = 128 pixel rows x BK, B = block_n rows x BK (both multiples of 1 KB)
    p.pair = (pair_enabled() && !p.halo && p.tiles_w * p.tiles_h * p.tiles_n >= 2) ? 1 : 0;
    p.a_bytes = 128u * p.bk * 2u;
    p.b_bytes = static_cast<uint32_t>(p.pair ? p.block_n / 2 : p.block_n) * p.bk * 2u;
    p.b_off = p.a_bytes;
    p.stage_stride = p.a_bytes + p.b_bytes;
    p.stages = std::max(1, std::min({kMaxStages, static_cast<int>(kSmemBudget / p.stage_stride), p.ntaps * p.kpt}));
    int smem_total = p.stages * static_cast<int>(p.stage_stride);
    if (p.halo) {
        // patch slots (18 x 16 pixels x 128 B) + weight ring; two CTAs per SM when it fits in 100 KB each
        p.a_bytes = 18u * 16u * 128u;
        p.na = p.kpt >= 2 ? 2 : 1;
        const int budget = (p.na * static_cast<int>(p.a_bytes) + 3 * static_cast<int>(p.b_bytes) <= kSmemBudget)
                               ? kSmemBudget : 2 * kSmemBudget;
        p.stages = std::max(2, std::min({kMaxStages, (budget - p.na * static_cast<int>(p.a_bytes)) / static_cast<int>(p.b_bytes),
                                         p.ntaps * p.kpt}));
        smem_total = p.na * static_cast<int>(p.a_bytes) + p.stages * static_cast<int>(p.b_bytes);
    }
    p.tmem_cols = p.block_n <= 32 ? 32u : p.block_n <= 64 ? 64u : 128u;
    p.acc_stride = static_cast<uint32_t>((p.block_n + 31) / 32 * 32);
    // split-K: a layer whose tiles cover less than half the SMs but whose K loop is long is cut along K;
    // each split keeps at least 3 k-blocks
    {
        const long ctas = static_cast<long>(p.tiles_w) * p.tiles_h * p.tiles_n * (d.cout_pad / p.block_n);
        const int num_it = p.ntaps * p.kpt;
        int splits = 1;
        if (split_k_enabled() && !p.halo && !p.pair && ctas * 3 <= 148 && num_it >= 16) {
            const int want = static_cast<int>(std::min<long>(16, 148 / ctas));
            const int ips = std::max(4, (num_it + want - 1) / want);
            splits = (num_it + ips - 1) / ips;
            p.it_per_split = ips;
        }
        if (splits <= 2) { splits = 1; p.it_per_split = num_it; }   // two-way splits do not pay for the extra sync
        p.splits = splits;
        p.stages = std::max(1, std::min(p.stages, p.it_per_split));
        p.part_ld = (p.block_n + 31) / 32 * 32;
        // slim variant: N <= 64 and more CTAs than two full waves of the wide variant
        p.slim = (slim_enabled() && !p.halo && !p.pair && splits == 1 && p.block_n <= 64 && ctas > 2 * 148) ? 1 : 0;
        if (p.slim) {
            p.stages = std::max(1, std::min({kMaxStages, static_cast<int>(kSmemBudgetSlim / p.stage_stride), num_it}));
            smem_total = p.stages * static_cast<int>(p.stage_stride);
        }
        // two MMA issue streams (wide variant, plain per-tap mode): even ring size, at least two k-blocks
        // (pays when the layer is one wave of CTAs, i.e. latency bound; multi-wave layers already overlap CTAs)
        p.dual = (dual_enabled() && !p.slim && !p.halo && !p.pair && splits == 1 && num_it >= 2 && p.stages >= 2 &&
                  ctas <= 2 * 148) ? 1 : 0;
        if (p.dual) {
            p.dual = 2;
            // one CTA per SM anyway: the whole shared memory can hold the ring (deeper pipeline per stream)
            // and four streams fit (their accumulators may then take all 512 TMEM columns)
            if (ctas <= 148) {
                p.stages = std::max(2, std::min({kMaxStages, static_cast<int>(2 * kSmemBudget / p.stage_stride), num_it}));
                if (quad_enabled() && p.stages >= 4 && num_it >= 4) p.dual = 4;
            }
            p.stages &= ~(p.dual - 1);
            smem_total = p.stages * static_cast<int>(p.stage_stride);
            const uint32_t need = p.dual * p.acc_stride;
            p.tmem_cols = need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
        }
    }
    const int out_align = d.out_f32 ? 4 : 8;
    p.vec_ok = (d.out_pitch % out_align == 0 && d.out_coff % out_align == 0 &&
                (d.res == nullptr || (d.res_pitch % 8 == 0 && d.res_coff % 8 == 0))) ? 1 : 0;
    const int S = d.stride;
    for (int r = 0; r < d.k; ++r)
        for (int s = 0; s < d.k; ++s) {
            int4 t;
            if (d.k == 1) t = make_int4(0, 0, 0, 0);
            else if (S == 1) t = make_int4(0, s - 1, 0, r - 1);
            else t = make_int4((s == 1 ? 0 : 1) * d.in_pitch, s == 0 ? -1 : 0, r == 1 ? 0 : 1, r == 0 ? -1 : 0);
            p.tap[r * d.k + s] = t;
        }
    p.out = d.out; p.out_pitch = d.out_pitch; p.out_coff = d.out_coff; p.out_f32 = d.out_f32;
    p.bias = d.bias; p.act = d.act;
    p.res = d.res; p.res_pitch = d.res_pitch; p.res_coff = d.res_coff;
    p.dup = d.dup; p.dup_pitch = d.dup_pitch; p.dup_coff = d.dup_coff; p.dup_mode = d.dup_mode;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A/B=f16 (0), K-major both,
    // N>>3 at [17,23), M>>4 at [24,29)
    p.idesc = (1u << 4) | (static_cast<uint32_t>(p.block_n >> 3) << 17) |
              (static_cast<uint32_t>((p.pair ? 256 : 128) >> 4) << 24);
    p.sbo = (p.bk == 64) ? 1024u : 512u;
    p.layout = (p.bk == 64) ? 2u : 4u;
    const CUtensorMapSwizzle swz = (p.bk == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

    if (p.halo) {
        // activation map for the halo patch: (Cpitch, W, H, N), box 64 ch x 16 x 18 x 1
        const cuuint64_t cp = static_cast<cuuint64_t>(d.in_pitch);
        cuuint64_t dims[4] = {cp, static_cast<cuuint64_t>(d.w_in), static_cast<cuuint64_t>(d.h_in), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[3] = {cp * 2, d.w_in * cp * 2, static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        cuuint32_t box[4] = {64u, 16u, 18u, 1u};
        encode(&l.tm_a, const_cast<__half*>(d.in), 4, dims, strides, box, swz);
    } else
    // activation map: (S*Cpitch, W/S, S, H/S, N)
    {
        const cuuint64_t cp = static_cast<cuuint64_t>(d.in_pitch);
        cuuint64_t dims[5] = {S * cp, static_cast<cuuint64_t>(d.w_in / S), static_cast<cuuint64_t>(S),
                              static_cast<cuuint64_t>(d.h_in / S), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[4] = {S * cp * 2, d.w_in * cp * 2, S * d.w_in * cp * 2,
                                 static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        cuuint32_t box[5] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.tw), 1u,
                             static_cast<cuuint32_t>(p.th), static_cast<cuuint32_t>(p.tn)};
        encode(&l.tm_a, const_cast<__half*>(d.in), 5, dims, strides, box, swz);
    }
    {
        const cuuint64_t ktot = static_cast<cuuint64_t>(p.ntaps) * d.cin_pad;
        cuuint64_t dims[2] = {ktot, static_cast<cuuint64_t>(d.cout_pad)};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.pair ? p.block_n / 2 : p.block_n)};
        encode(&l.tm_b, const_cast<__half*>(d.w), 2, dims, strides, box, swz);
    }
    l.grid = dim3(p.tiles_w * p.tiles_h * p.tiles_n, d.cout_pad / p.block_n, p.splits);
    if (p.pair) l.grid.x = (l.grid.x + 1) / 2 * 2;   // clusters of two M-tiles; a padding tile is all out of range
    l.smem_bytes = smem_total + 1024;
    l.flops = 2.0 * d.n * d.h_out * d.w_out * static_cast<double>(d.cout) * d.k * d.k * d.cin;
    return l;
}

static bool g_use_pdl = true;

// scratch of a split-K launch: one arrival counter per output tile (zero between launches), then the
// fp32 partial tiles [tile][split][128][part_ld]
size_t conv_scratch_bytes(const ConvLaunch& l) {
    if (l.v2) return conv2_scratch_bytes(l);
    if (l.p.splits <= 1) return 0;
    const size_t tiles = static_cast<size_t>(l.grid.x) * l.grid.y;
    return (tiles * sizeof(int) + 255) / 256 * 256 + tiles * l.p.splits * 128 * l.p.part_ld * sizeof(float);
}

void conv_bind_scratch(ConvLaunch& l, void* zeroed_base) {
    if (l.v2) { conv2_bind_scratch(l, zeroed_base); return; }
    if (l.p.splits <= 1) return;
    const size_t tiles = static_cast<size_t>(l.grid.x) * l.grid.y;
    l.p.counters = static_cast<int*>(zeroed_base);
    l.p.partial = reinterpret_cast<float*>(static_cast<char*>(zeroed_base) + (tiles * sizeof(int) + 255) / 256 * 256);
}


void conv_init() {
    conv2_init();
    // function attributes are per device (primary context): one flag per device ordinal
    static std::once_flag once_dev[64];
    int dev = 0;
    RMR_CUDA(cudaGetDevice(&dev));
    std::call_once(once_dev[dev & 63], [] {
        auto prep = [](auto kernel, int smem) {
            RMR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            RMR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        };
        prep(conv_umma_kernel<0, false, false>, 2 * kSmemBudget + 1024);
        prep(conv_umma_kernel<0, false, true>, 2 * kSmemBudget + 1024);
        prep(conv_umma_kernel<0, true, false>, kSmemBudgetSlim + 1024);
        prep(conv_umma_kernel<0, true, true>, kSmemBudgetSlim + 1024);
        prep(conv_umma_kernel<1>, kSmemBudget + 1024);
        prep(conv_umma_kernel<2>, 2 * kSmemBudget + 1024);
        prep(conv_umma_kernel<3>, kSmemBudget + 1024);
        get_encode_fn();
        const char* e = std::getenv("RMR_NO_PDL");
        g_use_pdl = !(e && e[0] == '1');
    });
}

void launch_conv_umma(const ConvLaunch& l, cudaStream_t s, bool pdl) {
    if (l.v2) { launch_conv2(l, s, pdl); return; }
    conv_init();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = l.grid;
    cfg.blockDim = dim3(l.p.slim ? kThreadsSlim : kThreads);
    cfg.dynamicSmemBytes = static_cast<size_t>(l.smem_bytes);
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_use_pdl && pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (l.p.pair) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const bool res = l.p.res != nullptr;
    if (l.p.pair) RMR_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<3>, l.tm_a, l.tm_b, l.p));
    else if (l.p.halo) RMR_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<2>, l.tm_a, l.tm_b, l.p));
    else if (l.p.splits > 1) RMR_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<1>, l.tm_a, l.tm_b, l.p));
    else if (l.p.slim && res) RMR_CUDA((cudaLaunchKernelEx(&cfg, conv_umma_kernel<0, true, true>, l.tm_a, l.tm_b, l.p)));
    else if (l.p.slim) RMR_CUDA((cudaLaunchKernelEx(&cfg, conv_umma_kernel<0, true, false>, l.tm_a, l.tm_b, l.p)));
    else if (res) RMR_CUDA((cudaLaunchKernelEx(&cfg, conv_umma_kernel<0, false, true>, l.tm_a, l.tm_b, l.p)));
    else RMR_CUDA((cudaLaunchKernelEx(&cfg, conv_umma_kernel<0, false, false>, l.tm_a, l.tm_b, l.p)));
}

void launch_conv_simt(const ConvDesc& d, cudaStream_t s) {
    static const bool stem_mma = [] { const char* e = std::getenv("RMR_STEM_SIMT"); return !(e && e[0] == '1'); }();
    if (stem_mma && d.cin_pad == 4 && d.k == 3 && d.stride == 2 && d.res == nullptr && !d.out_f32 && d.in_pitch == 4 &&
        d.in_coff == 0 && d.cout == 32 && d.cout_pad == 32 && d.out_pitch % 8 == 0 && d.out_coff % 8 == 0 && d.dup == nullptr) {
        const int tiles_w = (d.w_out + 15) / 16;
        const long total = static_cast<long>(d.n) * d.h_out * tiles_w;
        const int blocks = static_cast<int>(std::min<long>((total + 3) / 4, 148L * 12));
        conv_stem_mma_kernel<<<blocks, 128, 0, s>>>(d, tiles_w, total);
    } else if (d.cin_pad == 4 && d.k == 3 && d.stride == 2 && d.act == 1 && d.res == nullptr && !d.out_f32 &&
        d.in_pitch == 4 && d.in_coff == 0 && (d.cout == 32 || d.cout == 16)) {
        const long total = static_cast<long>(d.n) * d.h_out * ((d.w_out + 1) / 2);
        const int blocks = static_cast<int>((total + 127) / 128);
        if (d.cout == 32) conv_stem_kernel<32><<<blocks, 128, 0, s>>>(d);
        else conv_stem_kernel<16><<<blocks, 128, 0, s>>>(d);
    } else {
        const long total = static_cast<long>(d.n) * d.h_out * d.w_out * ((d.cout + 7) / 8);
        conv_simt_kernel<<<static_cast<int>((total + 127) / 128), 128, 0, s>>>(d);
    }
    RMR_CUDA(cudaGetLastError());
}

void launch_maxpool5(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff, int n,
                     int h, int w, int c, cudaStream_t s) {
    const long total = static_cast<long>(n) * h * w * (c / 8);
    maxpool5_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(in, in_pitch, in_coff, out, out_pitch,
                                                                          out_coff, n, h, w, c / 8);
    RMR_CUDA(cudaGetLastError());
}

void launch_sppf_pool3(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int coff1, int coff2,
                       int coff3, int n, int h, int w, int c, cudaStream_t s) {
    sppf_pool3_kernel<<<dim3(c / 8, n), 256, static_cast<size_t>(2) * h * w * sizeof(uint4), s>>>(
        in, in_pitch, in_coff, out, out_pitch, coff1, coff2, coff3, h, w);
    RMR_CUDA(cudaGetLastError());
}

void launch_upsample2(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff, int n,
                      int h_in, int w_in, int c, cudaStream_t s) {
    const long total = static_cast<long>(n) * h_in * 2 * w_in * 2 * (c / 8);
    upsample2_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(in, in_pitch, in_coff, out, out_pitch,
                                                                           out_coff, n, h_in, w_in, c / 8);
    RMR_CUDA(cudaGetLastError());
}

void launch_copy_channels(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff,
                          int n, int h, int w, int c, cudaStream_t s) {
    const long npix = static_cast<long>(n) * h * w;
    const long total = npix * (c / 8);
    copy_channels_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(in, in_pitch, in_coff, out,
                                                                               out_pitch, out_coff, npix, c / 8);
    RMR_CUDA(cudaGetLastError());
}

}  // namespace rmr
