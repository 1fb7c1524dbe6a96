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
constexpr bool kConv2Default = true;    // RMR_CONV_V2=0 falls back to the round-1 kernel (conv.cu)

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

struct DupMaps {
    CUtensorMap m[4];
};

// kHalo: 3x3 stride-1 layers, one [18][pw] pixel patch per channel chunk serves the nine taps.
// (tile_w, tile_h, tile_n) of m-tile j0, j0 + gm, j0 + 2 gm, ... without a division per tile: the step is decomposed
// once and added with carries (a 32-bit division by a run-time value is ~100 cycles; four of them per tile were
// a third of an epilogue-bound tile, profiles/r2_timeline_epi.txt)
struct TileWalk {
    int w, h, n, gw, gh, gn, tiles_w, tiles_h;
    // (gw, gh, gn) = the step gm in tile coordinates, decomposed on the host (Conv2Params::gm_w/h/n)
    __device__ __forceinline__ TileWalk(int j0, const Conv2Params& p) : gw(p.gm_w), gh(p.gm_h), gn(p.gm_n), tiles_w(p.tiles_w), tiles_h(p.tiles_h) {
        w = j0 % tiles_w; const int t = j0 / tiles_w; h = t % tiles_h; n = t / tiles_h;
    }
    __device__ __forceinline__ void next() {
        w += gw;
        int c = w >= tiles_w ? 1 : 0;
        w -= c * tiles_w;
        h += gh + c;
        c = h >= tiles_h ? 1 : 0;
        h -= c * tiles_h;
        n += gn + c;
    }
};

// kDbg: the instrumented build behind rmr_conv_timeline.  The stamps are compiled out of the product kernel: clock64()
// is a scheduling barrier even when predicated off, and a dozen of them in the epilogue cost 12 us (car) + 22 us (armor,
// 7 ROIs) per frame (profiles/r2_summary.md, "debug stamps").
template <bool kHalo, bool kRes, bool kDbg>
__global__ void __launch_bounds__(kThreads2, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_res,
             const __grid_constant__ DupMaps tm_dup, const __grid_constant__ Conv2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[kMaxA], bar_aempty[kMaxA];
    __shared__ __align__(8) uint64_t bar_bfull[kMaxB], bar_bempty[kMaxB];
    __shared__ __align__(8) uint64_t bar_accfull[2], bar_accempty[2], bar_bres;
    // TMA epilogue, one set per 32-column sub-tile of the staging buffer: written by the epilogue (outready),
    // read out by the TMA store (stfree), shortcut tile landed (resfull)
    __shared__ __align__(8) uint64_t bar_outready[8], bar_stfree[8], bar_resfull[8];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int4 s_tap[9];
    __shared__ float s_bias[256];

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // profiling aid: slot 0 start, 1 setup done, 2 + 8 t + {0: A producer past its empty wait, 1: issuer past the
    // accumulator wait, 2: first operands seen, 3: all MMAs of the tile issued, 4: epilogue sees the accumulator,
    // 5: accumulator read out, 6: first epilogue group stored, 7: second group stored} for tiles t < 7, 60 exit
    long long* dbg = (kDbg && p.dbg) ? p.dbg + static_cast<size_t>(blockIdx.x) * 64 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    // this CTA's slice: output channels [ch0, ch0 + block_n), pixel tiles j0, j0 + gm, ...
    const int nt = blockIdx.x % p.n_tiles;
    const int j0 = blockIdx.x / p.n_tiles;
    const int ch0 = nt * p.block_n;
    // the K loop in units: per-tap mode counts (tap, chunk) pairs tap-major (= k-blocks in weight order), halo mode
    // counts channel chunks (nine k-blocks each)
    const int u0 = 0;
    const int u1 = kHalo ? p.kpt : p.ntaps * p.kpt;

    if (warp == 1) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_a);
            tma_prefetch_desc(&tm_b);
            if (p.tma_epi) {
                tma_prefetch_desc(&tm_out);
                if (kRes && p.res != nullptr) tma_prefetch_desc(&tm_res);
                if (p.dup_mode) tma_prefetch_desc(&tm_dup.m[0]);
            }
        }
        if (lane < p.ntaps) s_tap[lane] = p.tap[lane];
    } else if (warp == 2) {
        // barrier init spread over the lanes (up to ~100 barriers: one thread would spend ~1500 cycles on them)
        for (int i = lane; i < p.sa; i += 32) {
            mbar_init(smem_u32(&bar_afull[i]), 1);
            mbar_init(smem_u32(&bar_aempty[i]), 1);
        }
        for (int i = lane; i < p.sb; i += 32) {
            mbar_init(smem_u32(&bar_bfull[i]), 1);
            mbar_init(smem_u32(&bar_bempty[i]), 1);
        }
        if (lane < 2) {
            mbar_init(smem_u32(&bar_accfull[lane]), 1);
            mbar_init(smem_u32(&bar_accempty[lane]), 8);   // one arrival per epilogue warp
        }
        if (lane == 2) mbar_init(smem_u32(&bar_bres), 1);
        if (p.tma_epi && lane >= 8 && lane < 8 + p.nchunks) {
            const int i = lane - 8;
            mbar_init(smem_u32(&bar_outready[i]), 128);   // the four warps (one per TMEM lane quarter) of a chunk
            mbar_init(smem_u32(&bar_stfree[i]), 1);
            mbar_init(smem_u32(&bar_resfull[i]), 1);
        }
        fence_barrier_init();
        __syncwarp();
        tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
        tmem_relinquish();
    }
    // bias: a constant, fetched now but parked in a register — the CTA-wide barrier below must not wait for a global
    // load (~700 cycles of the ~1600-cycle setup); the epilogue warps publish it among themselves later
    float bias_reg = 0.f;
    if (warp >= 4 && static_cast<int>(threadIdx.x) - 128 < p.block_n) bias_reg = __ldg(p.bias + ch0 + threadIdx.x - 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    pdl_launch2();
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    const uint32_t a_ring = smem_base + p.a_off;
    const uint32_t b_ring = smem_base + p.b_off;

    // The three issuing roles run warp-converged: every lane walks the loop and waits on the barriers, one
    // elected lane issues.  Issued from divergent code, ptxas wraps every UTMALDG / UTCHMMA / UTCBAR in an
    // ELECT + BRA.U.ANY lane loop (~150 cycles per instruction, tools/umma_bench.cu); with elect.sync under
    // warp-uniform control flow the bookkeeping stays in the uniform datapath and the MMAs issue back to back.
    // Stage granularity.  Every barrier round trip of the issuer (try_wait on a completed barrier ~150 cycles,
    // commit, loop) costs about as much as four N <= 64 MMAs (profiles/r2_timeline3.txt: 300 cycles for two waits
    // + 390 for four MMAs and a commit per k-block), so a stage carries several k-blocks:
    //   per-tap mode   g k-blocks of activations per stage (+ their weights when the weights stream: "joint" ring)
    //   halo mode      one patch = nine k-blocks; streamed weights come three taps per stage
    //   resident weights  the whole slice on ONE barrier, waited for once
    if (warp == 0 || warp == 1) {
        // ------------------------------ operand producers ------------------------------
        // warp 0: activations; warp 1: weights, and in per-tap mode every second activation box of a stage (one
        // UTMALDG keeps the issuing warp busy for ~340 cycles, profiles/r2_timeline3.txt: a k-block of N <= 128 MMAs
        // is shorter than that, so one warp cannot keep up).
        // Weights come through a 3-D view (bk, Cout, k-block): ONE instruction loads the CTA's whole resident slice,
        // or the g k-blocks of a joint stage, laid out k-block-major = one UMMA B tile after the other.
        // (slot, phase) ring bookkeeping without divisions; a fresh barrier passes a wait on parity 1.
        const bool leader = elect_one2();
        const bool is_a = warp == 0;
        const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
        const uint32_t aend = 8u * p.sa, a_stage = p.a_stage, a_kb = p.a_kb, b_kb = p.b_kb;
        const int bk = p.bk, kpt = p.kpt, g = p.g;
        if (!is_a && p.b_resident) {
            // weights are constants: no grid dependency, the slice is on its way while the previous layer still runs
            if (leader) {
                const uint32_t bres = smem_u32(&bar_bres);
                mbar_expect_tx(bres, static_cast<uint32_t>(p.ntaps * kpt) * b_kb);
                tma_load_3d(b_ring, &tm_b, bres, 0, ch0, 0);
            }
            __syncwarp();
        }
        if (kHalo) {
            if (is_a) {
                pdl_wait2();   // activations are the previous layer's output
                uint32_t aoff = 0, adst = a_ring, phase = 0;
                int ti = 0;
                TileWalk tile(j0, p);
                for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++ti) {
                    const int ow0 = tile.w * p.tw, oh0 = tile.h * p.th, n0 = tile.n * p.tn;
                    int cx = p.cin_coff;
                    for (int c = 0; c < kpt; ++c, cx += bk) {
                        mbar_wait_spin(aempty0 + aoff, phase ^ 1u);
                        if (dbg && leader && c == 0 && ti < 7 && !p.dbg_mode) dbg[2 + 8 * ti] = clock64();
                        if (leader) {
                            mbar_expect_tx(afull0 + aoff, p.a_tx);
                            if (p.halo == 2) {
                                // class (row parity pr, column parity pc): odd rows / columns start one pair earlier
#pragma unroll
                                for (int cls = 0; cls < 4; ++cls) {
                                    const int pr = cls >> 1, pc = cls & 1;
                                    tma_load_5d(adst + cls * p.cls_stride, &tm_a, afull0 + aoff, cx + pc * p.in_pitch, ow0 - pc, pr,
                                                oh0 - pr, n0);
                                }
                            } else {
                                tma_load_4d(adst, &tm_a, afull0 + aoff, cx, ow0 - 1, oh0 - 1, n0);
                            }
                        }
                        __syncwarp();
                        aoff += 8u; adst += a_stage;
                        if (aoff == aend) { aoff = 0; adst = a_ring; phase ^= 1u; }
                    }
                    if (mt + p.gm < p.m_tiles) tile.next();   // after this tile's loads are on their way
                }
            } else if (!p.b_resident) {
                // streamed weights, three taps (one kernel row) per stage; k-block of (tap, chunk c) = tap * kpt + c
                const uint32_t bfull0 = smem_u32(&bar_bfull[0]), bempty0 = smem_u32(&bar_bempty[0]);
                const uint32_t bend = 8u * p.sb, b_stage = p.b_stage;
                uint32_t boff = 0, bdst = b_ring, phase = 0;
                for (int mt = j0; mt < p.m_tiles; mt += p.gm) {
                    for (int c = 0; c < kpt; ++c) {
                        int kb = c;
                        for (int dy = 0; dy < 3; ++dy) {
                            mbar_wait_spin(bempty0 + boff, phase ^ 1u);
                            if (leader) {
                                mbar_expect_tx(bfull0 + boff, 3u * b_kb);
                                tma_load_3d(bdst, &tm_b, bfull0 + boff, 0, ch0, kb);
                                tma_load_3d(bdst + b_kb, &tm_b, bfull0 + boff, 0, ch0, kb + kpt);
                                tma_load_3d(bdst + 2u * b_kb, &tm_b, bfull0 + boff, 0, ch0, kb + 2 * kpt);
                            }
                            __syncwarp();
                            kb += 3 * kpt;
                            boff += 8u; bdst += b_stage;
                            if (boff == bend) { boff = 0; bdst = b_ring; phase ^= 1u; }
                        }
                    }
                }
            }
        } else {
            // per-tap mode: a stage = g k-blocks of activations (+ their weights when the weights stream).  Both warps walk
            // the same stages; warp 0 arms the barrier and issues the even boxes, warp 1 the weight box and the odd ones.
            const int nunits = p.ntaps * kpt;
            const uint32_t stage_tx_b = p.joint ? static_cast<uint32_t>(g) * b_kb : 0u;   // the weight box is always g k-blocks (zero fill past the end)
            const int parity = is_a ? 0 : 1;
            bool waited = is_a;
            if (is_a) pdl_wait2();   // activations are the previous layer's output
            uint32_t aoff = 0, adst = a_ring, phase = 0;
            int ti = 0;
            TileWalk tile(j0, p);
            for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++ti) {
                const int ow0 = tile.w * p.tw, oh0 = tile.h * p.th, n0 = tile.n * p.tn;
                int tap = 0, kc = 0;
                int4 tp = s_tap[0];
                int cx = p.cin_coff + tp.x;
                for (int u = 0; u < nunits; u += g) {
                    const int gg = min(g, nunits - u);
                    // warp 1 waits only for stages it puts something into (the weight box of a joint stage, the odd
                    // activation boxes).  A warp that only watches a barrier can miss a whole phase — the parity wait
                    // then never catches up and the warp is still waiting when the CTA is done (a hang at the end of
                    // the tile loop, seen once the debug stamps no longer slowed warp 0 down).  A stage that needs a
                    // box of warp 1 cannot complete twice without it, so its own waits never alias.
                    if (is_a || p.joint || gg > 1) mbar_wait_spin(aempty0 + aoff, phase ^ 1u);
                    if (dbg && leader && is_a && u == 0 && ti < 7 && !p.dbg_mode) dbg[2 + 8 * ti] = clock64();
                    if (leader) {
                        if (is_a) mbar_expect_tx(afull0 + aoff, static_cast<uint32_t>(gg) * a_kb + stage_tx_b);
                        else if (p.joint) tma_load_3d(adst + p.b_in_stage, &tm_b, afull0 + aoff, 0, ch0, u);
                    }
                    if (!waited && gg > 1) {   // warp 1's first activation box: from here on it reads the previous layer's output
                        pdl_wait2();
                        waited = true;
                    }
                    uint32_t dst = adst;
                    for (int j = 0; j < gg; ++j, dst += a_kb) {
                        if (leader && (j & 1) == parity) tma_load_5d(dst, &tm_a, afull0 + aoff, cx, ow0 + tp.y, tp.z, oh0 + tp.w, n0);
                        cx += bk;
                        if (++kc == kpt) {
                            kc = 0;
                            ++tap;
                            tp = s_tap[min(tap, 8)];
                            cx = p.cin_coff + tp.x;
                        }
                    }
                    __syncwarp();
                    aoff += 8u; adst += a_stage;
                    if (aoff == aend) { aoff = 0; adst = a_ring; phase ^= 1u; }
                }
                if (mt + p.gm < p.m_tiles) tile.next();   // after this tile's loads are on their way
            }
        }
    } else if (warp == 2) {
        // ------------------------------ MMA issuer ------------------------------
        // One warp feeds the tensor core, so the loop body is what bounds N <= 64 layers: everything a k-block
        // needs is kept in registers (no parameter loads, no divisions, no 64-bit descriptor rebuilds — a
        // descriptor is a constant high word and a low word that advances by a constant).
        // (No tcgen05.fence after the operand barriers: TMA's complete_tx on the mbarrier is what makes the tile
        // visible to the tensor core, both sides are the async proxy.  The fence stays where TMEM changes hands,
        // after the accumulator-empty wait.)
        const bool leader = elect_one2();
        const bool resident = p.b_resident != 0;
        const uint32_t idesc = p.idesc;
        const bool k4 = p.bk == 64;
        // descriptor words (umma_smem_desc): lo = addr >> 4 | LBO(1) << 16, hi = SBO >> 4 | version << 14 | layout << 29
        const uint32_t a_hi = (p.sbo_a >> 4) | (1u << 14) | (p.layout << 29);
        const uint32_t b_hi = (p.sbo_b >> 4) | (1u << 14) | (p.layout << 29);
        const uint32_t a_lo0 = (a_ring >> 4) | (1u << 16), b_lo0 = (b_ring >> 4) | (1u << 16);
        const uint32_t a_step = p.a_stage >> 4, b_step = p.b_stage >> 4;      // per stage
        const uint32_t a_kb_step = p.a_kb >> 4, b_kb_step = p.b_kb >> 4;      // per k-block
        const uint32_t b_in_stage = p.b_in_stage >> 4;
        const uint32_t tap_step = static_cast<uint32_t>(p.kpt) * b_kb_step;   // resident slice: k-blocks in weight order
        // operand tile of tap (dy, dx) inside a patch stage = row term + column term (16-byte units).  Stride 1: one
        // 18 x pw patch, the tap is a shifted window.  Stride 2: four parity-class patches of 17 x 9; the middle row /
        // column reads the even class, the outer ones the odd class, the far one shifted by one
        const uint32_t row16 = p.row_bytes >> 4, pitch16 = (static_cast<uint32_t>(p.pw) * p.row_bytes) >> 4, cls16 = p.cls_stride >> 4;
        const bool s2 = p.halo == 2;
        const uint32_t ty0 = s2 ? 2u * cls16 : 0u, ty1 = s2 ? 0u : pitch16, ty2 = s2 ? 2u * cls16 + pitch16 : 2u * pitch16;
        const uint32_t tx0 = s2 ? cls16 : 0u, tx1 = s2 ? 0u : row16, tx2 = s2 ? cls16 + row16 : 2u * row16;
        // barrier addresses: base + 8 * slot (the generic -> shared conversion is not free: once per role)
        const uint32_t afull0 = smem_u32(&bar_afull[0]), aempty0 = smem_u32(&bar_aempty[0]);
        const uint32_t bfull0 = smem_u32(&bar_bfull[0]), bempty0 = smem_u32(&bar_bempty[0]);
        const uint32_t accfull0 = smem_u32(&bar_accfull[0]), accempty0 = smem_u32(&bar_accempty[0]);
        const uint32_t aend = 8u * p.sa, bend = 8u * p.sb;
        const int g = p.g;
        auto mma_kblock = [&](uint32_t acc, uint32_t alo, uint32_t blo, uint32_t accumulate) {
            const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | alo;
            const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | blo;
            umma_f16(acc, ad, bd, idesc, accumulate);
            umma_f16(acc, ad + 2u, bd + 2u, idesc, 1u);
            if (k4) {
                umma_f16(acc, ad + 4u, bd + 4u, idesc, 1u);
                umma_f16(acc, ad + 6u, bd + 6u, idesc, 1u);
            }
        };
        uint32_t it = 0;
        uint32_t aoff = 0, boff = 0;            // 8 * slot
        uint32_t aphase = 0, bphase = 0;
        uint32_t a_lo = a_lo0, b_lo = b_lo0;
        if (resident) mbar_wait_spin(smem_u32(&bar_bres), 0u);   // the weight slice is in place
        for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++it) {
            const uint32_t ab = it & 1u;
            mbar_wait_spin(accempty0 + 8u * ab, ((it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            if (dbg && leader && it < 7 && p.dbg_mode != 2) dbg[3 + 8 * it] = clock64();
            const uint32_t acc = tmem_base + ab * p.acc_stride;
            uint32_t accumulate = 0;
            if (resident) b_lo = b_lo0;   // resident weights: slot = k-block index
            if (kHalo) {
                for (int c = u0; c < u1; ++c) {
                    mbar_wait_spin(afull0 + aoff, aphase);
                    if (dbg && leader && c == u0 && it < 7 && p.dbg_mode != 2) dbg[4 + 8 * it] = clock64();
                    if (resident) {
                        // nine k-blocks of MMAs and nothing else; the weight tile of (tap, chunk c) is k-block tap * kpt + c
                        if (leader) {
                            uint32_t w_lo = b_lo;
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy) {
                                const uint32_t row_lo = a_lo + (dy == 0 ? ty0 : dy == 1 ? ty1 : ty2);
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) {
                                    mma_kblock(acc, row_lo + (dx == 0 ? tx0 : dx == 1 ? tx1 : tx2), w_lo, accumulate);
                                    accumulate = 1;
                                    w_lo += tap_step;
                                }
                            }
                            umma_commit(aempty0 + aoff);   // patch consumed
                        }
                        __syncwarp();
                        accumulate = 1;
                        b_lo += b_kb_step;   // next chunk
                    } else {
#pragma unroll 1
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint32_t row_lo = a_lo + (dy == 0 ? ty0 : dy == 1 ? ty1 : ty2);
                            mbar_wait_spin(bfull0 + boff, bphase);   // the three taps of this kernel row
                            if (leader) {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) {
                                    mma_kblock(acc, row_lo + (dx == 0 ? tx0 : dx == 1 ? tx1 : tx2), b_lo + dx * b_kb_step, accumulate);
                                    accumulate = 1;
                                }
                                umma_commit(bempty0 + boff);
                                if (dy == 2) umma_commit(aempty0 + aoff);   // patch consumed
                            }
                            __syncwarp();
                            accumulate = 1;
                            b_lo += b_step;
                            boff += 8u;
                            if (boff == bend) { boff = 0; bphase ^= 1u; b_lo = b_lo0; }
                        }
                    }
                    a_lo += a_step;
                    aoff += 8u;
                    if (aoff == aend) { aoff = 0; aphase ^= 1u; a_lo = a_lo0; }
                }
            } else {
#pragma unroll 1
                for (int u = u0; u < u1; u += g) {
                    const int gg = min(g, u1 - u);
                    mbar_wait_spin(afull0 + aoff, aphase);
                    if (dbg && leader && u == u0 && it < 7 && p.dbg_mode != 2) dbg[4 + 8 * it] = clock64();
                    if (leader) {
                        uint32_t alo = a_lo, blo = resident ? b_lo : a_lo + b_in_stage;
                        for (int j = 0; j < gg; ++j) {
                            mma_kblock(acc, alo, blo, accumulate);
                            accumulate = 1;
                            alo += a_kb_step;
                            blo += b_kb_step;
                        }
                        umma_commit(aempty0 + aoff);
                    }
                    __syncwarp();
                    accumulate = 1;
                    b_lo += static_cast<uint32_t>(gg) * b_kb_step;
                    a_lo += a_step;
                    aoff += 8u;
                    if (aoff == aend) { aoff = 0; aphase ^= 1u; a_lo = a_lo0; }
                }
            }
            if (leader) umma_commit(accfull0 + 8u * ab);   // accumulator of this tile complete
            if (dbg && leader && it < 7 && p.dbg_mode != 2) dbg[5 + 8 * it] = clock64();
            __syncwarp();
        }
    } else if (warp == 3) {
        // ------------------------------ epilogue DMA (TMA epilogue only) ------------------------------
        // Shortcut tiles in, output tiles out, one 32-column sub-tile at a time.  The epilogue warps never touch
        // global memory: a row-per-thread tile moved with 16-byte accesses costs one L1 request per thread and
        // instruction (16 N cycles per tile for the stores alone, profiles/r2_timeline_v2.txt); through shared
        // memory it is 2 N cycles of conflict-free STS and one bulk store the TMA unit turns into full lines.
        if (p.tma_epi) {
            const bool leader = elect_one2();
            const bool has_res = kRes && p.res != nullptr;
            const uint32_t stage = smem_base + p.stage_off;
            const uint32_t outready0 = smem_u32(&bar_outready[0]), stfree0 = smem_u32(&bar_stfree[0]), resfull0 = smem_u32(&bar_resfull[0]);
            const uint32_t res_bytes = 128u * 64u;   // [128 rows][32 fp16]
            pdl_wait2();   // shortcut reads and output writes must follow the previous grid
            TileWalk tile(j0, p);
            int ow0 = tile.w * p.tw, oh0 = tile.h * p.th, n0 = tile.n * p.tn;
            if (has_res && j0 < p.m_tiles) {
                if (leader)
                    for (int sidx = 0; sidx < p.nchunks; ++sidx) {
                        mbar_expect_tx(resfull0 + 8u * sidx, res_bytes);
                        tma_load_4d(stage + sidx * p.sub_bytes, &tm_res, resfull0 + 8u * sidx, p.res_coff + ch0 + 32 * sidx, ow0, oh0, n0);
                    }
                __syncwarp();
            }
            uint32_t it = 0;
            for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++it) {
                const bool more = mt + p.gm < p.m_tiles;
                if (more) tile.next();   // the tile after this one: its shortcut is fetched as soon as a sub-tile is free
                const int nw0 = tile.w * p.tw, nh0 = tile.h * p.th, nn0 = tile.n * p.tn;
                for (int sidx = 0; sidx < p.nchunks; ++sidx) {
                    mbar_wait_spin(outready0 + 8u * sidx, it & 1u);
                    if (dbg && leader && sidx == 0 && it < 7 && p.dbg_mode == 2) dbg[9 + 8 * it] = clock64();
                    if (leader) {
                        const uint32_t src = stage + sidx * p.sub_bytes;
                        tma_store_4d(&tm_out, src, p.out_coff + ch0 + 32 * sidx, ow0, oh0, n0);
                        if (p.dup_mode == 1) {
                            tma_store_4d(&tm_dup.m[0], src, p.dup_coff + ch0 + 32 * sidx, ow0, oh0, n0);
                        } else if (p.dup_mode == 2) {
                            // nearest 2x upsample written by the producer: phase (dy, dx) of the destination is a strided
                            // view with the source's extents, so the same box lands on pixels (2 oh + dy, 2 ow + dx)
#pragma unroll
                            for (int ph = 0; ph < 4; ++ph) tma_store_4d(&tm_dup.m[ph], src, p.dup_coff + ch0 + 32 * sidx, ow0, oh0, n0);
                        }
                        bulk_commit_group();
                        bulk_wait_read0();   // the sub-tile has been read out: it may be refilled
                        if (has_res) {
                            if (more) {
                                mbar_expect_tx(resfull0 + 8u * sidx, res_bytes);
                                tma_load_4d(src, &tm_res, resfull0 + 8u * sidx, p.res_coff + ch0 + 32 * sidx, nw0, nh0, nn0);
                            }
                        } else {
                            mbar_arrive(stfree0 + 8u * sidx);
                        }
                    }
                    __syncwarp();
                }
                ow0 = nw0; oh0 = nh0; n0 = nn0;
            }
            // the global writes are complete before the CTA retires.  (Leaving them to the grid's own completion —
            // RMR_EXIT_WAIT_ALL=0, what CUTLASS's TMA-store epilogues do — measured no faster: 752.3 vs 751.6 frames/s.)
            if (leader && p.exit_wait_all) bulk_wait_all();
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
        if (static_cast<int>(threadIdx.x) - 128 < p.block_n) s_bias[threadIdx.x - 128] = bias_reg;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps only
        if (!p.tma_epi) pdl_wait2();   // residual reads and output writes must follow the previous grid
        uint32_t it = 0;
        for (int mt = j0; mt < p.m_tiles; mt += p.gm, ++it) {
            // pixel of this thread's row: only the direct-store epilogue addresses global memory itself (four integer
            // divisions per tile — ~500 cycles the TMA epilogue, whose tiles are addressed by warp 3, must not pay)
            int ow = 0, oh = 0, n = 0;
            bool valid = false;
            size_t pix = 0;
            const __half* rptr = nullptr;
            if (!p.tma_epi) {
                int t = mt;
                const int tile_w = t % p.tiles_w; t /= p.tiles_w;
                const int tile_h = t % p.tiles_h;
                const int tile_n = t / p.tiles_h;
                ow = tile_w * p.tw + tw_i; oh = tile_h * p.th + th_i; n = tile_n * p.tn + tn_i;
                valid = (ow < p.w_out) && (oh < p.h_out) && (n < p.n);
                pix = (static_cast<size_t>(n) * p.h_out + oh) * p.w_out + ow;
                rptr = (kRes && p.res != nullptr && valid) ? p.res + pix * p.res_pitch + p.res_coff + ch0 : nullptr;
            }
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

            const bool fine = dbg && threadIdx.x == 128 && it < 7 && p.dbg_mode == 2;   // epilogue steps of warp 4
            if (fine) dbg[2 + 8 * it] = clock64();
            mbar_wait_spin(smem_u32(&bar_accfull[ab]), (it >> 1) & 1);
            tc_fence_after();
            if (dbg && threadIdx.x == 128 && it < 7 && !p.dbg_mode) dbg[6 + 8 * it] = clock64();
            if (fine) dbg[3 + 8 * it] = clock64();
            // the last chunk this warp reads: once it is in registers the accumulator goes back to the issuer
            int last_c0 = -1;
            for (int c0 = chunk0; c0 < nvalid; c0 += 64) last_c0 = c0;
            if (p.tma_epi) {
                // ---- TMA epilogue: accumulator chunk -> registers -> (+ shortcut from the staged tile) -> staging
                // sub-tile in the swizzled layout the bulk store expects -> warp 3 stores it ----
                const bool has_res = kRes && p.res != nullptr;
                const uint32_t stage = smem_base + p.stage_off;
                const bool f32out = p.out_f32 != 0;
                // 16-byte slot j of row `row` sits at j ^ swz: SWIZZLE_64B rows (fp16, 64 B) / SWIZZLE_128B rows (fp32, 128 B)
                const uint32_t swz = f32out ? static_cast<uint32_t>(row & 7) : static_cast<uint32_t>((row >> 1) & 3);
                const uint32_t row_off = static_cast<uint32_t>(row) * p.chunk_bytes;
                for (int c0 = chunk0; c0 < nvalid; c0 += 64) {
                    const int sidx = c0 >> 5;
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32(taddr + c0, v);
                    tmem_ld_wait();
                    if (fine && c0 == chunk0) dbg[4 + 8 * it] = clock64();
                    if (c0 == last_c0) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                        if (dbg && threadIdx.x == 128 && it < 7 && !p.dbg_mode) dbg[7 + 8 * it] = clock64();
                    }
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float h = 0.5f * f[j];
                            f[j] = fmaf(h, tanh_approx(h), h);
                        }
                    } else if (p.act == 2) {
#pragma unroll
                        for (int h0 = 0; h0 < 32; h0 += 16) {
                            float e[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) e[j] = 1.0f + ex2_approx(-1.4426950408889634f * f[h0 + j]);
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[h0 + j] = f[h0 + j] * rcp_approx2(e[j]);
                        }
                    }
                    const uint32_t sub = stage + sidx * p.sub_bytes + row_off;
                    if (has_res) {
                        mbar_wait_spin(smem_u32(&bar_resfull[sidx]), it & 1u);   // the shortcut tile landed (and the sub-tile is ours)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 r;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                                         : "r"(sub + ((static_cast<uint32_t>(j) ^ swz) << 4)));
                            const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 a = __half22float2(h[u]);
                                f[8 * j + 2 * u] += a.x; f[8 * j + 2 * u + 1] += a.y;
                            }
                        }
                    } else {
                        if (fine && c0 == chunk0) dbg[5 + 8 * it] = clock64();
                        mbar_wait_spin(smem_u32(&bar_stfree[sidx]), (it & 1u) ^ 1u);   // the previous tile's store has read it out
                    }
                    if (fine && c0 == chunk0) dbg[6 + 8 * it] = clock64();
                    if (f32out) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sub + ((static_cast<uint32_t>(j) ^ swz) << 4)),
                                         "f"(f[4 * j]), "f"(f[4 * j + 1]), "f"(f[4 * j + 2]), "f"(f[4 * j + 3]) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const __half2 hh = __floats2half2_rn(f[8 * j + 2 * u], f[8 * j + 2 * u + 1]);
                                w[u] = *reinterpret_cast<const uint32_t*>(&hh);
                            }
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sub + ((static_cast<uint32_t>(j) ^ swz) << 4)),
                                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                        }
                    }
                    if (fine && c0 == chunk0) dbg[7 + 8 * it] = clock64();
                    fence_proxy_async();   // generic-proxy writes -> visible to the bulk store
                    mbar_arrive(smem_u32(&bar_outready[sidx]));
                    if (fine && c0 == chunk0) dbg[8 + 8 * it] = clock64();
                }
                if (dbg && it < 7 && !p.dbg_mode) {
                    if (threadIdx.x == 128) dbg[8 + 8 * it] = clock64();
                    if (threadIdx.x == 256) dbg[9 + 8 * it] = clock64();
                }
                if (last_c0 < 0) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                }
            } else {
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
                        if (dbg && threadIdx.x == 128 && it < 7 && !p.dbg_mode) dbg[7 + 8 * it] = clock64();
                    }
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    if (valid) finish_chunk(c0, f);
                }
                if (dbg && it < 7 && !p.dbg_mode) {
                    if (threadIdx.x == 128) dbg[8 + 8 * it] = clock64();
                    if (threadIdx.x == 256) dbg[9 + 8 * it] = clock64();
                }
                if (last_c0 < 0) {   // this warp has no chunk (N <= 32 and warp >= 8): still release the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bar_accempty[ab]));
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[60] = clock64();
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
             const cuuint32_t* box, CUtensorMapSwizzle swz, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_fn2()(tm, dtype, rank, base, dims, strides_bytes, box, estr,
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
    // stride 2: the nine taps of an 8 x 16 output tile read four parity classes of the input (even / odd rows x even /
    // odd columns), each a dense 9 x 17 pixel patch in the (pixel pair, parity) view the per-tap mode already uses —
    // four boxes per channel chunk instead of nine strided ones, half the rows (p.halo == 2)
    // Only single-chunk layers (Cin <= 64) take it: there the accumulation order (tap by tap) is the per-tap mode's, so
    // the results are bit-identical; with several chunks the order becomes chunk-major, which moved one near-tie of the
    // armor NMS on asset frame 1 (two candidates 0.7010 / 0.7014) for a gain of < 1 % of the step.
    const bool halo2_ok = d.k == 3 && S == 2 && p.kpt == 1 && d.h_out >= halo_min && d.w_out >= halo_min && d.w_in % 2 == 0 &&
                          d.h_in % 2 == 0;
    p.halo = (env_flag("RMR_HALO", true) && halo_ok) ? 1 : (env_flag("RMR_HALO_S2", true) && halo2_ok) ? 2 : 0;
    if (p.halo == 1) {
        p.tw = 8; p.th = 16; p.tn = 1;
        p.pw = env_int("RMR_HALO_PW", 10);
    } else if (p.halo == 2) {
        p.tw = 8; p.th = 16; p.tn = 1;
        p.pw = 9;
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
    const int force_splits = 0;
    // TMA epilogue (shared-memory staged tile, bulk store): whole 32-channel chunks, 16-byte aligned views, one
    // destination or a plain second copy; the 2x-upsample destination and ragged channel counts keep the direct stores
    const int out_align0 = d.out_f32 ? 4 : 8;
    const bool views_ok = d.out_pitch % out_align0 == 0 && d.out_coff % out_align0 == 0 &&
                          (d.res == nullptr || (d.res_pitch % 8 == 0 && d.res_coff % 8 == 0 && !d.out_f32)) &&
                          (d.dup == nullptr || (d.dup_pitch % 8 == 0 && d.dup_coff % 8 == 0 && !d.out_f32));
    const bool tma_epi_ok = env_flag("RMR_TMA_EPI", true) && views_ok && d.cout % 32 == 0 && d.cout == d.cout_pad && force_splits <= 1;
    const int esize = d.out_f32 ? 4 : 2;
    const int kSM = 148;
    double best_cost = -1;
    int best_n = 0, best_splits = 1;
    for (int bn = 256; bn >= 16; bn -= 16) {
        if (d.cout_pad % bn != 0) continue;
        if (force_n && bn != force_n && d.cout_pad % force_n == 0) continue;
        if (tma_epi_ok && bn % 32 != 0) continue;
        bool single_patch_slot = false;
        if (p.halo == 2) {
            // one 4-class patch stage + two weight stages of three taps (or the resident slice) + the staging tile must
            // fit; with room for one patch slot only, loading a patch and multiplying it take turns
            const long a_stage2 = (4L * ((17L * 9L * p.row_bytes + 127L) & ~127L) + 1023L) & ~1023L;
            const long b_stream = 2L * 3L * bn * p.bk * 2L, b_res = 9L * p.kpt * bn * p.bk * 2L;
            const long stg = tma_epi_ok ? 128L * bn * esize : 0L;
            const long room = kSmemMax - 1024 - stg - std::min(b_stream, b_res);
            if (a_stage2 > room) continue;
            single_patch_slot = 2 * a_stage2 > room;
        }
        const int n_tiles = d.cout_pad / bn;
        // split-K (deterministic: fp32 partial tiles through L2, last-arriving split reduces) stays available for
        // experiments (RMR_CONV_SPLITS=n) but is not planned: a partial tile is 128 x N x 4 B written and read
        // once per split — for the layers that would want it that is 10-30x the layer's own traffic (measured:
        // 40x40x256 -> 384 stride 2 at batch 7, 8 splits: 228 us against 25 us for the round-1 kernel).  Narrower
        // channel tiles give the small maps their parallelism instead: below N = 64 an MMA costs the same ~56
        // issue cycles whatever N is, so more, narrower CTAs are free until the activation re-reads show.
        const int max_splits = 1;
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
            const double a_write = p.halo == 1 ? 4096.0 * (18.0 * p.pw) / (9.0 * 128.0)
                                   : p.halo == 2 ? 4096.0 * (4.0 * 17.0 * 9.0) / (9.0 * 128.0) : 4096.0;
            const double staging = tma_epi_ok ? 128.0 * bn * esize : 0.0;
            const bool resident = static_cast<double>(kb) * bn * p.bk * 2 + (p.halo ? 2.0 : 3.0) * 16384 + staging <= kSmemMax - 1024 &&
                                  (p.halo || tiles_per_cta > 1);
            const double b_write = resident ? bn * 32.0 / tiles_per_cta : bn * 32.0;
            const double smem_cyc = (4096.0 + bn * 32.0 + a_write + b_write) / 128.0;
            const double ingest_rate = std::min(125.0, 11000.0 / busy);
            const double slice = std::max({56.0, bn / 2.0, smem_cyc, (a_write + b_write) / ingest_rate});
            const double per_kb = ksteps * slice;
            // epilogue of one tile: 8 warps, 32-column chunks (TMEM load, bias, SiLU, fp16 stores); overlaps the next
            // tile's main loop, so a tile costs the larger of the two
            // (direct stores: one L1 request per thread and 16 bytes, 16 N cycles for an fp16 tile)
            const double epi = (tma_epi_ok ? 500.0 + bn * 4.0 : 400.0 + bn * 8.0 * esize * (d.res ? 2.0 : 1.0)) +
                               (splits > 1 ? 4000.0 + bn * 30.0 * splits : 0.0);
            const double tile_cyc = std::max(kb * per_kb + (single_patch_slot ? ups * 1600.0 : 0.0), epi);
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
    static const bool exit_wait_all = env_flag("RMR_EXIT_WAIT_ALL", true);
    p.exit_wait_all = exit_wait_all ? 1 : 0;
    p.gm_w = p.gm % p.tiles_w;
    p.gm_h = (p.gm / p.tiles_w) % p.tiles_h;
    p.gm_n = p.gm / p.tiles_w / p.tiles_h;
    const int kb_cta = p.units_per_split * kb_per_unit;

    // ---- shared-memory layout: [A (or joint) ring][weights: resident slice | halo stream ring][epilogue staging] ----
    p.b_kb = static_cast<uint32_t>(p.block_n) * p.bk * 2u;
    if (p.halo == 1) {
        p.a_tx = 18u * p.pw * p.row_bytes;
        p.a_kb = (p.a_tx + 1023u) & ~1023u;
    } else if (p.halo == 2) {
        // one parity class: 17 rows of 9 pixels.  Class bases are 128-byte aligned only (RMR_HALO_S2_ALIGN=1024 pads them
        // to the swizzle period): TMA and UMMA both derive the swizzle phase from the absolute shared-memory address, and
        // the 3.5 KB of padding per stage is what decides between one and two stages for the N = 64 layers
        const uint32_t cls_tx = 17u * 9u * p.row_bytes;
        const uint32_t al = static_cast<uint32_t>(env_int("RMR_HALO_S2_ALIGN", 128));
        p.cls_stride = (cls_tx + al - 1u) / al * al;
        p.a_tx = 4u * cls_tx;
        p.a_kb = (4u * p.cls_stride + 1023u) & ~1023u;
    } else {
        p.a_tx = 128u * p.row_bytes;
        p.a_kb = p.a_tx;
    }
    p.tma_epi = (tma_epi_ok && p.splits == 1) ? 1 : 0;
    p.esize_out = esize;
    p.nchunks = p.block_n / 32;
    p.chunk_bytes = 32u * esize;
    p.sub_bytes = 128u * p.chunk_bytes;
    const int stage_bytes = p.tma_epi ? static_cast<int>(p.sub_bytes) * p.nchunks : 0;
    const int tiles_per_cta = (p.m_tiles + p.gm - 1) / p.gm;
    // (Capping single-tile CTAs at half an SM of shared memory, so that the next layer's CTA can become resident under
    // programmatic dependent launch and run its prologue during this layer's tail, was measured and lost: the shallower
    // ring costs more per k-block than the overlap returns — car 0.53 vs 0.44 ms, armor 0.94 vs 0.79 ms.)
    const long budget = kSmemMax - 1024 - stage_bytes;
    const long a_units_total = static_cast<long>(tiles_per_cta) * p.units_per_split;   // A units (patches / k-blocks) this CTA ever loads
    const long b_slice = static_cast<long>(kb_cta) * p.b_kb;
    const bool want_resident = env_flag("RMR_B_RESIDENT", true);
    long b_region = 0;
    p.g = 1; p.joint = 0; p.b_in_stage = 0; p.sb = 0; p.b_stage = p.b_kb;
    if (p.halo) {
        p.a_stage = p.a_kb;
        p.b_resident = (want_resident && b_slice + std::min<long>(a_units_total, 2) * p.a_stage <= budget) ? 1 : 0;
        if (p.b_resident) {
            b_region = b_slice;
        } else {
            // streamed weights: three taps (one kernel row) per stage; two or three patch slots, the rest is weight stages
            p.b_stage = 3u * p.b_kb;
            const long a_res = std::min<long>(a_units_total, 3) * p.a_stage;
            p.sb = static_cast<int>(std::max<long>(2, std::min<long>({kMaxB, (budget - a_res) / p.b_stage, 3 * a_units_total})));
            b_region = static_cast<long>(p.sb) * p.b_stage;
        }
        p.sa = static_cast<int>(std::max<long>(1, std::min<long>({kMaxA, a_units_total, (budget - b_region) / p.a_stage})));
    } else {
        // a CTA with one tile has nothing to reuse: its weights stream in the activation stages (joint ring, one
        // barrier per stage); with several tiles the slice stays resident when it fits
        p.b_resident = (want_resident && tiles_per_cta > 1 && b_slice + 3L * p.a_kb <= budget) ? 1 : 0;
        const long kb_bytes = p.a_kb + (p.b_resident ? 0 : p.b_kb);
        b_region = p.b_resident ? b_slice : 0;
        p.joint = p.b_resident ? 0 : 1;
        // k-blocks per stage: as many as leave three stages in flight, at most four
        p.g = static_cast<int>(std::max<long>(1, std::min<long>({4, kb_cta, (budget - b_region) / (3 * kb_bytes)})));
        if (const int fg = env_int("RMR_CONV_G", 0)) p.g = std::max(1, std::min(fg, kb_cta));
        p.a_stage = static_cast<uint32_t>(p.g * kb_bytes);
        p.b_in_stage = static_cast<uint32_t>(p.g) * p.a_kb;
        const long stages_total = static_cast<long>(tiles_per_cta) * ((kb_cta + p.g - 1) / p.g);
        p.sa = static_cast<int>(std::max<long>(1, std::min<long>({kMaxA, stages_total, (budget - b_region) / p.a_stage})));
    }
    if (static_cast<long>(p.sa) * p.a_stage + b_region > budget)
        throw CudaError("conv2: operand ring does not fit in shared memory");
    p.a_off = 0;
    p.b_off = static_cast<uint32_t>(p.sa) * p.a_stage;
    p.stage_off = p.b_off + static_cast<uint32_t>(b_region);   // 1 KB aligned: every stage is a multiple of 1 KB
    l.smem_bytes = static_cast<int>(p.stage_off) + stage_bytes + 1024;

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
    p.in_pitch = d.in_pitch;
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
    if (p.halo == 1) {
        cuuint64_t dims[4] = {cp, static_cast<cuuint64_t>(d.w_in), static_cast<cuuint64_t>(d.h_in), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[3] = {cp * 2, d.w_in * cp * 2, static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.pw), 18u, 1u};
        encode2(&l.tm_a, const_cast<__half*>(d.in), 4, dims, strides, box, swz);
    } else {
        cuuint64_t dims[5] = {S * cp, static_cast<cuuint64_t>(d.w_in / S), static_cast<cuuint64_t>(S),
                              static_cast<cuuint64_t>(d.h_in / S), static_cast<cuuint64_t>(d.n)};
        cuuint64_t strides[4] = {S * cp * 2, d.w_in * cp * 2, S * d.w_in * cp * 2, static_cast<cuuint64_t>(d.h_in) * d.w_in * cp * 2};
        // (channels of S adjacent pixels, pixel pairs, row parity, row pairs, image): a per-tap box is one parity of
        // tw x th pixel pairs; a stride-2 halo box is one parity class of 9 x 17
        cuuint32_t box[5] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.halo == 2 ? 9 : p.tw), 1u,
                             static_cast<cuuint32_t>(p.halo == 2 ? 17 : p.th), static_cast<cuuint32_t>(p.tn)};
        encode2(&l.tm_a, const_cast<__half*>(d.in), 5, dims, strides, box, swz);
    }
    {
        // weights [Cout_pad][K] as (bk, Cout_pad, K / bk): k-block outermost, so a box of several k-blocks lands as
        // consecutive [N][bk] UMMA tiles.  Box depth: the whole slice (resident), g k-blocks (joint ring) or one
        const cuuint64_t ktot = static_cast<cuuint64_t>(p.ntaps) * d.cin_pad;
        const int kb_total = p.ntaps * p.kpt;
        const int depth = p.b_resident ? kb_total : (p.halo ? 1 : p.g);
        cuuint64_t dims[3] = {static_cast<cuuint64_t>(p.bk), static_cast<cuuint64_t>(d.cout_pad), static_cast<cuuint64_t>(kb_total)};
        cuuint64_t strides[2] = {ktot * 2, static_cast<cuuint64_t>(p.bk) * 2};
        cuuint32_t box[3] = {static_cast<cuuint32_t>(p.bk), static_cast<cuuint32_t>(p.block_n), static_cast<cuuint32_t>(depth)};
        encode2(&l.tm_b, const_cast<__half*>(d.w), 3, dims, strides, box, swz);
    }
    if (p.tma_epi) {
        // output / shortcut / second destination: (C, W, H, N) views, box = one 32-channel sub-tile of the pixel tile
        auto view = [&](CUtensorMap* tm, void* base, int pitch, bool f32) {
            const cuuint64_t es = f32 ? 4 : 2, cpi = static_cast<cuuint64_t>(pitch);
            cuuint64_t dims[4] = {cpi, static_cast<cuuint64_t>(d.w_out), static_cast<cuuint64_t>(d.h_out), static_cast<cuuint64_t>(d.n)};
            cuuint64_t strides[3] = {cpi * es, d.w_out * cpi * es, static_cast<cuuint64_t>(d.h_out) * d.w_out * cpi * es};
            cuuint32_t box[4] = {32u, static_cast<cuuint32_t>(p.tw), static_cast<cuuint32_t>(p.th), static_cast<cuuint32_t>(p.tn)};
            encode2(tm, base, 4, dims, strides, box, f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
        };
        view(&l.tm_out, d.out, d.out_pitch, d.out_f32 != 0);
        if (d.res) view(&l.tm_res, const_cast<__half*>(d.res), d.res_pitch, false);
        if (d.dup && d.dup_mode == 1) view(&l.tm_dup[0], d.dup, d.dup_pitch, false);
        if (d.dup && d.dup_mode == 2) {
            // destination [N][2H][2W][pitch]: phase (dy, dx) = base + (dy * 2W + dx) pixels, pixel steps doubled
            const cuuint64_t cpi = static_cast<cuuint64_t>(d.dup_pitch);
            for (int ph = 0; ph < 4; ++ph) {
                const int dy = ph >> 1, dx = ph & 1;
                __half* base = d.dup + (static_cast<size_t>(dy) * 2 * d.w_out + dx) * d.dup_pitch;
                cuuint64_t dims[4] = {cpi, static_cast<cuuint64_t>(d.w_out), static_cast<cuuint64_t>(d.h_out), static_cast<cuuint64_t>(d.n)};
                cuuint64_t strides[3] = {2 * cpi * 2, 2 * (2 * static_cast<cuuint64_t>(d.w_out)) * cpi * 2,
                                         (2 * static_cast<cuuint64_t>(d.h_out)) * (2 * d.w_out) * cpi * 2};
                cuuint32_t box[4] = {32u, static_cast<cuuint32_t>(p.tw), static_cast<cuuint32_t>(p.th), static_cast<cuuint32_t>(p.tn)};
                encode2(&l.tm_dup[ph], base, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B);
            }
        }
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
        prep(conv2_kernel<false, false, false>);
        prep(conv2_kernel<false, true, false>);
        prep(conv2_kernel<true, false, false>);
        prep(conv2_kernel<true, true, false>);
        prep(conv2_kernel<false, false, true>);
        prep(conv2_kernel<false, true, true>);
        prep(conv2_kernel<true, false, true>);
        prep(conv2_kernel<true, true, true>);
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
    DupMaps dm;
    std::memcpy(dm.m, l.tm_dup, sizeof(dm.m));
    auto go = [&](auto kernel) { RMR_CUDA((cudaLaunchKernelEx(&cfg, kernel, l.tm_a, l.tm_b, l.tm_out, l.tm_res, dm, l.q))); };
    if (l.q.dbg) {      // instrumented build (rmr_conv_timeline only)
        if (l.q.halo) { if (res) go(conv2_kernel<true, true, true>); else go(conv2_kernel<true, false, true>); }
        else { if (res) go(conv2_kernel<false, true, true>); else go(conv2_kernel<false, false, true>); }
    } else {
        if (l.q.halo) { if (res) go(conv2_kernel<true, true, false>); else go(conv2_kernel<true, false, false>); }
        else { if (res) go(conv2_kernel<false, true, false>); else go(conv2_kernel<false, false, false>); }
    }
}

}  // namespace rmr
