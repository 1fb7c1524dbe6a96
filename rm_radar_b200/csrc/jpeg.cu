// Baseline JPEG decoding on the device (the step before the hot path).
//
// The reference reads every camera frame with cv::imread (/root/reference/samples/main.cpp:24-40): libjpeg-turbo
// on one CPU core (tens of ms for its 2592x2048 frames), then 16 MB of BGR go over PCIe.  Here the file image
// (1 MB) is uploaded as it is and decoded by kernels into the BGR frame the detector reads:
//
//   unstuff   FF00 -> FF, RSTn markers removed, start bit of every restart interval   (count / scan / scatter)
//   entropy   one cooperative kernel.  The bit stream is cut into subsequences of `sub_bits`; thread i decodes
//             subsequence i from a state (bit position, zig-zag index, block-in-MCU) and hands its end state to
//             thread i+1.  States start as guesses and are iterated to the fixed point s[i+1] = f_i(s[i]) with
//             s[0] exact, which by induction is the sequential decode; Huffman codes self-synchronise, so a few
//             rounds suffice and only threads whose input changed redo their work.  A scan of the blocks
//             completed per subsequence gives every thread its output position; a last pass writes coefficients.
//   dc scan   DC differences -> predictors per MCU (segmented at restart intervals)
//   idct      dequantise + the 13-bit fixed-point LLM inverse DCT (jidctint.c, JDCT_ISLOW), one thread per block
//   colour    triangle-filter chroma upsampling (jdsample.c h2v1 / h2v2 "fancy") + YCbCr -> BGR (jdcolor.c)
//
// Integer arithmetic throughout: bit-exact with oracle/jpeg_ref.c, which is pinned to cv2.imdecode.
#include "jpeg.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace rmr {

namespace {

constexpr int kUnstuffThreads = 256, kUnstuffBytes = 16;            // bytes per thread
constexpr int kDecodeThreads = 64;
constexpr int kMaxRoundsSlack = 16;
enum : int { CTRL_BARRIER = 0, CTRL_STATUS = 1, CTRL_ROUNDS = 2, CTRL_BYTES = 3, CTRL_MARKERS = 4, CTRL_BLOCKS = 5,
             CTRL_DECODES = 6, CTRL_CHANGED = 8 };
enum : int { ST_BLOCK_COUNT = 1, ST_INTERVALS = 2, ST_TAIL = 4 };

constexpr int kLutBits = 10;
// lookup entry: bits 0-5 bits consumed (code + magnitude), 6-12 zig-zag advance (64 = end of block), 13-17 code length,
// 18-21 magnitude size s, 22-25 run r; 0 = the code is longer than kLutBits
__host__ __device__ inline uint32_t lut_entry(int len, int sym, bool dc) {
    const int s = sym & 15, r = dc ? 0 : sym >> 4;
    const int zinc = dc ? 1 : (s ? r + 1 : (r == 15 ? 16 : 64));
    return static_cast<uint32_t>(len + s) | (static_cast<uint32_t>(zinc) << 6) | (static_cast<uint32_t>(len) << 13) |
           (static_cast<uint32_t>(s) << 18) | (static_cast<uint32_t>(r) << 22);
}
constexpr int kLut2 = 512;            // second level: the 16-bit prefixes at and above the first code longer than kLutBits
struct JpegTables {
    uint32_t lut[6][1 << kLutBits];   // [component * 2 + ac][next 10 bits]
    uint32_t lut2[6][kLut2];          // [table][16-bit prefix - base16[table]]
    uint32_t base16[8];
    int maxcode[6][18];       // largest code of each length, -1 when the length is unused
    int valoff[6][18];        // index of the first symbol of the length minus its first code
    uint8_t vals[6][256];
    uint16_t quant[3][64];    // natural order
    uint8_t comp_of_block[8];
    uint8_t first_block_of_comp[4];
};


struct JpegDev {
    const uint8_t* raw;
    uint32_t n_raw;
    uint8_t* stream;
    int2* blk_counts;
    int2* blk_offsets;
    int n_ublocks;
    uint32_t* intervals;
    int n_intervals;
    uint2* states;
    uint2* cand;
    uint4* res;
    unsigned char* bmap;
    uint4* ck;                    // [n_sub][kLanes][kCheckpoints]: (bit position, zig-zag | block << 8, blocks so far)
    unsigned long long* tstamp;   // phase boundaries of the entropy kernel (block 0), globaltimer ns
    int n_sub;
    uint32_t sub_bits;
    int* ctrl;
    int* grid_sums;
    int16_t* coef;
    int16_t* dc_abs;
    int* mcu_dc;
    long n_blocks;
    int n_mcus, bpm, ncomp, restart;
    const JpegTables* tables;
    int width, height, mcus_x, mcus_y, hs, vs;
    uint8_t* plane[3];
    int pitch[3];
    uint8_t* bgr;
    int stride;
};

// ---------------------------------------------------------------------------------------------------------
// unstuffing
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_rst(unsigned b) { return (b & 0xF8u) == 0xD0u; }

// keep / marker bit masks of the 16 bytes starting at `base`
__device__ __forceinline__ void classify16(const uint8_t* __restrict__ raw, uint32_t n, uint32_t base, uint32_t& keep,
                                           uint32_t& mark, uint8_t (&b)[18]) {
    const uint4 v = *reinterpret_cast<const uint4*>(raw + base);   // the buffer is padded past n
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    b[0] = base ? raw[base - 1] : 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j + 1] = (w[j >> 2] >> (8 * (j & 3))) & 0xFF;
    b[17] = base + 16 < n ? raw[base + 16] : 0;
    keep = mark = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (base + j >= n) break;
        const unsigned cur = b[j + 1], prev = b[j], next = (base + j + 1 < n) ? b[j + 2] : 0u;
        const bool m = cur == 0xFF && is_rst(next);
        const bool drop = m || (prev == 0xFF && (cur == 0 || is_rst(cur)));
        if (m) mark |= 1u << j;
        if (!drop) keep |= 1u << j;
    }
}

__global__ void __launch_bounds__(kUnstuffThreads) jpeg_unstuff_count_kernel(JpegDev J) {
    const uint32_t base = (blockIdx.x * kUnstuffThreads + threadIdx.x) * kUnstuffBytes;
    uint32_t keep = 0, mark = 0;
    uint8_t b[18];
    if (base < J.n_raw) classify16(J.raw, J.n_raw, base, keep, mark, b);
    __shared__ int sk[kUnstuffThreads / 32], sm[kUnstuffThreads / 32];
    int k = __popc(keep), m = __popc(mark);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k += __shfl_xor_sync(0xffffffffu, k, o);
        m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    if ((threadIdx.x & 31) == 0) { sk[threadIdx.x >> 5] = k; sm[threadIdx.x >> 5] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tk = 0, tm = 0;
        for (int i = 0; i < kUnstuffThreads / 32; ++i) { tk += sk[i]; tm += sm[i]; }
        J.blk_counts[blockIdx.x] = make_int2(tk, tm);
    }
}

// exclusive scan of the per-block (bytes, markers) totals by one block; also closes the interval table
__global__ void __launch_bounds__(1024) jpeg_unstuff_scan_kernel(JpegDev J) {
    __shared__ int2 part[1024];
    const int per = (J.n_ublocks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, J.n_ublocks);
    int2 s = make_int2(0, 0);
    for (int i = lo; i < hi; ++i) { const int2 c = J.blk_counts[i]; s.x += c.x; s.y += c.y; }
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int2 t = make_int2(0, 0);
        if (threadIdx.x >= o) t = part[threadIdx.x - o];
        __syncthreads();
        part[threadIdx.x].x += t.x;
        part[threadIdx.x].y += t.y;
        __syncthreads();
    }
    int2 run = threadIdx.x ? part[threadIdx.x - 1] : make_int2(0, 0);
    for (int i = lo; i < hi; ++i) {
        const int2 c = J.blk_counts[i];
        J.blk_offsets[i] = run;
        run.x += c.x;
        run.y += c.y;
    }
    if (threadIdx.x == 1023) {
        const int2 tot = part[1023];
        J.ctrl[CTRL_BYTES] = tot.x;
        J.ctrl[CTRL_MARKERS] = tot.y;
        J.intervals[0] = 0;
        if (tot.y + 1 == J.n_intervals) {
            J.intervals[J.n_intervals] = static_cast<uint32_t>(tot.x) * 8u;
        } else {
            // wrong number of restart markers: decode as one interval, flag the frame
            atomicOr(&J.ctrl[CTRL_STATUS], ST_INTERVALS);
            for (int i = 1; i <= J.n_intervals; ++i) J.intervals[i] = static_cast<uint32_t>(tot.x) * 8u;
        }
    }
}

__global__ void __launch_bounds__(kUnstuffThreads) jpeg_unstuff_scatter_kernel(JpegDev J) {
    const uint32_t base = (blockIdx.x * kUnstuffThreads + threadIdx.x) * kUnstuffBytes;
    uint32_t keep = 0, mark = 0;
    uint8_t b[18];
    if (base < J.n_raw) classify16(J.raw, J.n_raw, base, keep, mark, b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int k = __popc(keep), m = __popc(mark);
    int ik = k, im = m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tk = __shfl_up_sync(0xffffffffu, ik, o), tm = __shfl_up_sync(0xffffffffu, im, o);
        if (lane >= o) { ik += tk; im += tm; }
    }
    __shared__ int wk[kUnstuffThreads / 32], wm[kUnstuffThreads / 32];
    if (lane == 31) { wk[warp] = ik; wm[warp] = im; }
    __syncthreads();
    const int2 boff = J.blk_offsets[blockIdx.x];
    int off = boff.x + ik - k, moff = boff.y + im - m;
    for (int w = 0; w < warp; ++w) { off += wk[w]; moff += wm[w]; }
    const bool markers_ok = J.ctrl[CTRL_MARKERS] + 1 == J.n_intervals;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if ((mark >> j) & 1u) {
            if (markers_ok) J.intervals[1 + moff] = static_cast<uint32_t>(off) * 8u;
            ++moff;
        }
        if ((keep >> j) & 1u) J.stream[off++] = b[j + 1];
    }
}

// ---------------------------------------------------------------------------------------------------------
// entropy decoding
// ---------------------------------------------------------------------------------------------------------
struct SmemTables {
    uint32_t lut[6][1 << kLutBits];
    uint32_t lut2[6][kLut2];
    uint32_t base16[8];
    int maxcode[6][18];
    int valoff[6][18];
    uint8_t vals[6][256];
};

__device__ __forceinline__ uint32_t peek32(const uint32_t* __restrict__ words, uint32_t p) {
    const uint32_t wi = p >> 5;
    const uint32_t a = __byte_perm(__ldg(words + wi), 0, 0x0123), b = __byte_perm(__ldg(words + wi + 1), 0, 0x0123);
    return __funnelshift_l(b, a, p & 31);
}

__device__ __forceinline__ void grid_barrier(int* counter, unsigned nblocks) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(counter), 1u);
        const unsigned target = (ticket / nblocks + 1u) * nblocks;
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
        __threadfence();
    }
    __syncthreads();
}

// Decodes from state `st` while the bit position is below `limit`; returns the end state.  state.x = bit position of
// the next code, state.y = zig-zag index of the next coefficient | block-in-MCU << 8.  Defined for ANY start state
// (speculative starts land inside codes): unknown codes consume 16 bits, run-aways end the block.
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// kCk: the state at the first code boundary at or after ck_next, ck_next + ck_step, ... (kCheckpoints of them) is recorded
// with the blocks completed so far, so that the write pass can split the subsequence over the lanes of its group.
constexpr int kCheckpoints = 8;
template <bool kWrite, bool kCk = false>
__device__ __forceinline__ uint2 decode_range(const JpegDev& J, const SmemTables& T, uint2 st, uint32_t limit, int& n_done,
                                              long blk0, uint4* ck_out = nullptr, uint32_t ck_next = 0, uint32_t ck_step = 0) {
    const uint32_t* words = reinterpret_cast<const uint32_t*>(J.stream);
    uint32_t p = st.x;
    int z = st.y & 0xFF, c = st.y >> 8;
    int ckj = 0;
    if (p >= limit) {
        if (kCk)
            for (; ckj < kCheckpoints; ++ckj) __stcg(ck_out + ckj, make_uint4(p, st.y, 0u, 0u));
        return st;
    }
    int k = 0;
    if (J.n_intervals > 1) {   // the restart interval p lies in
        int lo = 0, hi = J.n_intervals - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (J.intervals[mid] <= p) lo = mid; else hi = mid - 1;
        }
        k = lo;
    }
    uint32_t ivl_end = J.intervals[k + 1];
    // component of block c of an MCU without a table lookup: blocks 0 .. bpm-3 are luma, then Cb, Cr
    const int bpm = J.bpm, cshift = J.ncomp == 3 ? bpm - 3 : 1 << 20;
    int comp = max(0, c - cshift);
    uint32_t lut_base = smem_u32(&T.lut[0][0]);
    asm volatile("" : "+r"(lut_base));   // keep the shared-window addresses in registers (no S2R in the loop)
    constexpr uint32_t kAcOff = 4u << kLutBits;             // bytes per table: the AC table follows the DC table of a component
    uint32_t toff = (static_cast<uint32_t>(comp) << (kLutBits + 3)) | (z ? kAcOff : 0u);
    const bool multi = J.n_intervals > 1;
    // write pass only: block and MCU counters of the block being filled
    const int n_blocks = static_cast<int>(J.n_blocks);
    int blk = static_cast<int>(blk0), mcu = kWrite ? static_cast<int>(blk0 / bpm) : 0;
    int16_t* cptr = J.coef + (static_cast<size_t>(blk) << 6);
    // 96-bit window over the big-endian stream: w0:w1 cover bits [32 wi, 32 wi + 64); r2 is the (not yet byte-swapped)
    // word after them, loaded one refill ahead so that no load sits on the per-symbol dependency chain
    uint32_t wi = p >> 5;
    uint32_t w0 = __byte_perm(__ldg(words + wi), 0, 0x0123), w1 = __byte_perm(__ldg(words + wi + 1), 0, 0x0123),
             r2 = __ldg(words + wi + 2);
    uint32_t lut2_base = smem_u32(&T.lut2[0][0]);
    asm volatile("" : "+r"(lut2_base));
    while (p < limit) {
        if (kCk) {
            while (ckj < kCheckpoints && p >= ck_next) {
                __stcg(ck_out + ckj, make_uint4(p, static_cast<uint32_t>(z) | (static_cast<uint32_t>(c) << 8),
                                                static_cast<uint32_t>(n_done), 0u));
                ++ckj;
                ck_next += ck_step;
            }
        }
        uint32_t off = p - (wi << 5);
        if (off >= 32) {               // at most one word per symbol (a symbol is <= 31 bits); jumps reload below
            ++wi;
            w0 = w1;
            w1 = __byte_perm(r2, 0, 0x0123);
            r2 = __ldg(words + wi + 2);
            off -= 32;
        }
        const uint32_t w = __funnelshift_l(w1, w0, off);
        uint32_t e = lds_u32(lut_base + toff + ((w >> (30 - kLutBits)) & ((4u << kLutBits) - 4u)));
        if (e == 0) {                                         // a code longer than kLutBits
            const int t = static_cast<int>(toff >> (kLutBits + 2));
            const uint32_t i2 = (w >> 16) - T.base16[t];
            if (i2 < static_cast<uint32_t>(kLut2)) e = lds_u32(lut2_base + (static_cast<uint32_t>(t) * kLut2 + i2) * 4u);
            if (e == 0) {                                     // beyond the second level (or no such code): canonical search
                int len = 16, sym = 0;
#pragma unroll 1
                for (int l = kLutBits + 1; l <= 16; ++l) {
                    const int code = static_cast<int>(w >> (32 - l));
                    if (code <= T.maxcode[t][l]) {
                        len = l;
                        sym = T.vals[t][(code + T.valoff[t][l]) & 0xFF];
                        break;
                    }
                }
                e = lut_entry(len, sym, z == 0);
            }
        }
        p += e & 63u;
        if (kWrite) {
            const int s = (e >> 18) & 15, len = (e >> 13) & 31, idx = z + static_cast<int>((e >> 22) & 15u);
            if ((z == 0 || s) && idx < 64 && blk < n_blocks) {
                const uint32_t v = __funnelshift_l(w << len, 0u, s);            // the s bits after the code (0 when s = 0)
                const int val = static_cast<int>(v) - (v < ((1u << s) >> 1) ? (1 << s) - 1 : 0);
                cptr[idx] = static_cast<int16_t>(val);       // zig-zag position; DC: the difference, resolved by the DC scan
                if (z == 0 && val) atomicAdd(&J.mcu_dc[comp * J.n_mcus + mcu], val);
            }
        }
        z += static_cast<int>((e >> 6) & 127u);
        const bool done = z >= 64;
        z = done ? 0 : z;
        n_done += done ? 1 : 0;
        const int cn = c + 1 == bpm ? 0 : c + 1;
        c = done ? cn : c;
        comp = max(0, c - cshift);
        toff = (static_cast<uint32_t>(comp) << (kLutBits + 3)) | (done ? 0u : kAcOff);
        if (kWrite) {
            blk += done ? 1 : 0;
            cptr += done ? 64 : 0;
            mcu += (done && c == 0) ? 1 : 0;
        }
        if (done && c == 0 && (multi || p + 8 > ivl_end)) {
            // end of an MCU: fewer than 8 (padding) bits left in the interval -> continue at the next one
            while (k + 1 < J.n_intervals && J.intervals[k + 1] <= p) ++k;
            ivl_end = J.intervals[k + 1];
            const uint32_t rem = ivl_end > p ? ivl_end - p : 0u;
            if (rem < 8) {
                const bool pad = rem == 0 || bpm >= 3 || (peek32(words, p) >> (32 - rem)) == ((1u << rem) - 1u);
                if (pad) {
                    p = max(p, ivl_end);
                    if (k + 1 < J.n_intervals) { ++k; ivl_end = J.intervals[k + 1]; }
                    wi = p >> 5;                               // the jump may skip words: reload the window
                    w0 = __byte_perm(__ldg(words + wi), 0, 0x0123);
                    w1 = __byte_perm(__ldg(words + wi + 1), 0, 0x0123);
                    r2 = __ldg(words + wi + 2);
                }
            }
        }
    }
    if (kCk)      // checkpoints the subsequence ended before: empty tails
        for (; ckj < kCheckpoints; ++ckj)
            __stcg(ck_out + ckj, make_uint4(p, static_cast<uint32_t>(z) | (static_cast<uint32_t>(c) << 8), static_cast<uint32_t>(n_done), 0u));
    return make_uint2(p, static_cast<uint32_t>(z) | (static_cast<uint32_t>(c) << 8));
}

__global__ void __launch_bounds__(kDecodeThreads) jpeg_entropy_simple_kernel(JpegDev J) {
    __shared__ SmemTables T;
    __shared__ int s_scan[kDecodeThreads];
    __shared__ int s_any;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(J.tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        // lut / maxcode / valoff / vals are the leading members of both structs, in the same order
        constexpr int kWords = (sizeof(T.lut) + sizeof(T.lut2) + sizeof(T.base16) + sizeof(T.maxcode) + sizeof(T.valoff) + sizeof(T.vals)) / 4;
        for (int i = threadIdx.x; i < kWords; i += kDecodeThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int i = blockIdx.x * kDecodeThreads + threadIdx.x;
    const uint32_t total_bits = static_cast<uint32_t>(J.ctrl[CTRL_BYTES]) * 8u;
    // the grid is sized for the stuffed byte count; subsequences past the end of the unstuffed stream do not exist
    const int n_sub = static_cast<int>((static_cast<unsigned long long>(total_bits) + J.sub_bits - 1) / J.sub_bits);
    const bool active = i < n_sub;
    const uint32_t lo = static_cast<uint32_t>(i) * J.sub_bits;
    const uint32_t hi = active ? static_cast<uint32_t>(min(static_cast<unsigned long long>(lo) + J.sub_bits,
                                                           static_cast<unsigned long long>(total_bits))) : 0u;
    uint2 my_in = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu), my_out = make_uint2(0, 0);
    int my_n = 0;
    int round = 0;
    const int stride = J.n_sub + 1;
    for (;;) {
        const uint2* cur = J.states + (round & 1) * stride;
        uint2* nxt = J.states + ((round + 1) & 1) * stride;
        bool redo = false;
        if (active) {
            uint2 in;
            if (round == 0) in = make_uint2(lo, 0u);                 // guess: a block of component 0 starts here (exact for i = 0)
            else in = i == 0 ? make_uint2(0u, 0u) : __ldcg(cur + i);
            if (in.x != my_in.x || in.y != my_in.y) {
                my_in = in;
                my_n = 0;
                my_out = decode_range<false>(J, T, in, hi, my_n, 0);
                redo = true;
            }
            __stcg(nxt + i + 1, my_out);
        }
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        if (redo) s_any = 1;
        __syncthreads();
        if (threadIdx.x == 0 && s_any) atomicAdd(&J.ctrl[CTRL_CHANGED + round], 1);
        grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
        const int changed = *reinterpret_cast<volatile int*>(&J.ctrl[CTRL_CHANGED + round]);
        ++round;
        if (changed == 0 || round >= n_sub + kMaxRoundsSlack - 1) break;
    }
    // output position of every subsequence: exclusive scan of the blocks each one completes
    s_scan[threadIdx.x] = my_n;
    __syncthreads();
    for (int o = 1; o < kDecodeThreads; o <<= 1) {
        int t = 0;
        if (threadIdx.x >= o) t = s_scan[threadIdx.x - o];
        __syncthreads();
        s_scan[threadIdx.x] += t;
        __syncthreads();
    }
    if (threadIdx.x == kDecodeThreads - 1) __stcg(&J.grid_sums[blockIdx.x], s_scan[threadIdx.x]);
    grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
    long before = 0;
    for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += kDecodeThreads) before += __ldcg(&J.grid_sums[b]);
    __shared__ long s_before[kDecodeThreads];
    s_before[threadIdx.x] = before;
    __syncthreads();
    if (threadIdx.x == 0) {
        long t = 0;
        for (int j = 0; j < kDecodeThreads; ++j) t += s_before[j];
        s_before[0] = t;
    }
    __syncthreads();
    const long blk0 = s_before[0] + s_scan[threadIdx.x] - my_n;
    if (active) {
        int n2 = 0;
        decode_range<true>(J, T, my_in, hi, n2, blk0);
        if (i == n_sub - 1) {
            const long total = blk0 + my_n;
            J.ctrl[CTRL_BLOCKS] = static_cast<int>(total);
            J.ctrl[CTRL_ROUNDS] = round;
            if (total != J.n_blocks) atomicOr(&J.ctrl[CTRL_STATUS], ST_BLOCK_COUNT);
            if (my_out.y != 0) atomicOr(&J.ctrl[CTRL_STATUS], ST_TAIL);
        }
    }
}

// The same fixed point reached in ~4 sequential passes instead of ~17.  A wrong guess survives because the block-in-MCU
// index never self-synchronises, so every subsequence is decoded under all `bpm` hypotheses at once (8 lanes per
// subsequence, lane = block-in-MCU guess):
//   pass 0   lane k of subsequence i decodes from (first bit of i, DC expected, block k)  -> candidate states of boundary i+1
//   pass 1   lane k decodes subsequence i from candidate k of boundary i (a real state, whoever produced it) -> result,
//            and the index of that result among the candidates of boundary i+1 (the link; the true chain finds its
//            successor there for ~99 % of the boundaries)
//   chase    links are followed from (boundary 0, lane 0) in shared memory: per block over its 32 subsequences for every
//            entry lane, then over the per-block maps, then again inside the block from its true entry lane.  A dead
//            link is continued at lane 0 of the next boundary (a placeholder the verify loop replaces).
//   verify   the fixed-point loop of the simple kernel, seeded with the chased (input, output, count) triples.  A thread
//            whose input changes first looks the new input up among its candidates (pass-1 results are real decodes of
//            this subsequence) and only decodes when it is not there.  Every triple a thread holds is a genuine
//            f_i(input), so the loop still ends exactly at the sequential decode.
constexpr int kLanes = 8, kSubsPerBlock = 32, kHypThreads = kLanes * kSubsPerBlock, kMaxHypBlocks = 592;
constexpr uint32_t kInvalid = 0xFFFFFFFFu;
__device__ __forceinline__ void stamp(const JpegDev& J, int slot) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        J.tstamp[slot] = t;
    }
}
constexpr unsigned kDead = 0x80u;

__global__ void __launch_bounds__(kHypThreads) jpeg_entropy_kernel(JpegDev J) {
    __shared__ SmemTables T;
    __shared__ int s_scan[kHypThreads];
    __shared__ int s_any;
    __shared__ unsigned char s_link[kSubsPerBlock][kLanes];
    __shared__ unsigned char s_sel[kSubsPerBlock];
    __shared__ unsigned char s_bmap[kMaxHypBlocks][kLanes];
    __shared__ long s_before[kHypThreads / 32];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(J.tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&T);
        constexpr int kWords = (sizeof(T.lut) + sizeof(T.lut2) + sizeof(T.base16) + sizeof(T.maxcode) + sizeof(T.valoff) + sizeof(T.vals)) / 4;
        for (int i = threadIdx.x; i < kWords; i += kHypThreads) dst[i] = src[i];
    }
    __syncthreads();
    const int sub = threadIdx.x >> 3, lane = threadIdx.x & 7;
    const int i = blockIdx.x * kSubsPerBlock + sub;
    const uint32_t total_bits = static_cast<uint32_t>(J.ctrl[CTRL_BYTES]) * 8u;
    const int n_sub = static_cast<int>((static_cast<unsigned long long>(total_bits) + J.sub_bits - 1) / J.sub_bits);
    const bool active = i < n_sub, owner = active && lane == 0, hyp = active && lane < J.bpm;
    const uint32_t lo = static_cast<uint32_t>(i) * J.sub_bits;
    const uint32_t hi = active ? static_cast<uint32_t>(min(static_cast<unsigned long long>(lo) + J.sub_bits,
                                                           static_cast<unsigned long long>(total_bits))) : 0u;
    uint2* cand = J.cand;            // [n_sub + 1][kLanes]
    uint4* res = J.res;              // [n_sub][kLanes]: end state, blocks completed, link

    stamp(J, 0);
    // ---- pass 0: candidates of every boundary ----
    if (active) {
        uint2 e = make_uint2(kInvalid, kInvalid);
        if (hyp) {
            int n = 0;
            e = decode_range<false>(J, T, make_uint2(lo, static_cast<uint32_t>(lane) << 8), hi, n, 0);
        }
        __stcg(cand + static_cast<size_t>(i + 1) * kLanes + lane, e);
        if (i == 0) __stcg(cand + lane, lane == 0 ? make_uint2(0u, 0u) : make_uint2(kInvalid, kInvalid));
    }
    grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
    stamp(J, 1);

    // ---- pass 1: decode from every candidate, link the result to the candidates of the next boundary ----
    unsigned link = kDead;
    if (hyp) {
        const uint2 in = __ldcg(cand + static_cast<size_t>(i) * kLanes + lane);
        uint2 x = make_uint2(kInvalid, kInvalid);
        int n = 0;
        if (in.x != kInvalid) {
            x = decode_range<false, true>(J, T, in, hi, n, 0, J.ck + (static_cast<size_t>(i) * kLanes + lane) * kCheckpoints, lo,
                                          J.sub_bits / kCheckpoints);
            if (i + 1 < n_sub) {
                for (int k = 0; k < J.bpm; ++k) {
                    const uint2 c = __ldcg(cand + static_cast<size_t>(i + 1) * kLanes + k);
                    if (c.x == x.x && c.y == x.y) { link = static_cast<unsigned>(k); break; }
                }
            } else {
                link = 0;
            }
        }
        __stcg(res + static_cast<size_t>(i) * kLanes + lane, make_uint4(x.x, x.y, static_cast<uint32_t>(n), link));
    }
    s_link[sub][lane] = static_cast<unsigned char>(link);
    __syncthreads();

    // ---- chase, level 1: exit lane of this block for every entry lane ----
    if (threadIdx.x < kLanes) {
        unsigned k = threadIdx.x, dead = 0;
        for (int s2 = 0; s2 < kSubsPerBlock; ++s2) {
            if (static_cast<int>(blockIdx.x) * kSubsPerBlock + s2 >= n_sub) break;
            const unsigned l = s_link[s2][k];
            dead = l & kDead;
            k = dead ? 0u : l;
        }
        J.bmap[static_cast<size_t>(blockIdx.x) * kLanes + threadIdx.x] = static_cast<unsigned char>(k | dead);
    }
    grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
    stamp(J, 2);
    // ---- level 2: entry lane of this block = the maps of all blocks before it applied to lane 0 ----
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(J.bmap);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_bmap[0][0]);
        for (int w = threadIdx.x; w < static_cast<int>(blockIdx.x) * (kLanes / 4); w += kHypThreads) dst[w] = __ldcg(src + w);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned k = 0, dead = 0;
        for (int b = 0; b < static_cast<int>(blockIdx.x); ++b) {
            const unsigned m = s_bmap[b][k];
            dead = m & kDead;
            k = m & 7u;
        }
        // ---- level 3: true lane of every subsequence of this block ----
        (void)dead;
        for (int s2 = 0; s2 < kSubsPerBlock; ++s2) {
            s_sel[s2] = static_cast<unsigned char>(k);      // after a dead link: lane 0 as a placeholder (a genuine triple)
            const unsigned l = s_link[s2][k];
            k = (l & kDead) ? 0u : l;
        }
    }
    __syncthreads();

    // ---- seed the fixed-point loop with the chased triples ----
    uint2 my_in = make_uint2(kInvalid, kInvalid), my_out = make_uint2(kInvalid, kInvalid);
    int my_n = 0, my_slot = -1;       // my_slot: whose checkpoints describe the decode from my_in (lane k of pass 1, 6 = my own)
    const int stride = J.n_sub + 1;
    if (owner) {
        const unsigned sel = s_sel[sub] & 7u;
        const uint2 c = __ldcg(cand + static_cast<size_t>(i) * kLanes + sel);
        const uint4 r = __ldcg(res + static_cast<size_t>(i) * kLanes + sel);
        if (c.x != kInvalid) {
            my_in = c;
            my_out = make_uint2(r.x, r.y);
            my_n = static_cast<int>(r.z);
            my_slot = static_cast<int>(sel);
        }
        __stcg(J.states + stride + i + 1, my_out);       // read as `cur` by round 1
    }
    grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
    stamp(J, 3);
    int round = 1, decodes = 0;
    for (;;) {
        const uint2* cur = J.states + (round & 1) * stride;
        uint2* nxt = J.states + ((round + 1) & 1) * stride;
        bool redo = false;
        if (owner) {
            const uint2 in = i == 0 ? make_uint2(0u, 0u) : __ldcg(cur + i);
            if (in.x != my_in.x || in.y != my_in.y) {
                my_in = in;
                redo = true;
                bool found = false;
                if (in.x != kInvalid) {
                    for (int k = 0; k < J.bpm && !found; ++k) {
                        const uint2 c = __ldcg(cand + static_cast<size_t>(i) * kLanes + k);
                        if (c.x == in.x && c.y == in.y) {
                            const uint4 r = __ldcg(res + static_cast<size_t>(i) * kLanes + k);
                            my_out = make_uint2(r.x, r.y);
                            my_n = static_cast<int>(r.z);
                            my_slot = k;
                            found = true;
                        }
                    }
                }
                if (!found) {
                    my_n = 0;
                    my_slot = 6;
                    my_out = decode_range<false, true>(J, T, in, hi, my_n, 0, J.ck + (static_cast<size_t>(i) * kLanes + 6) * kCheckpoints,
                                                       lo, J.sub_bits / kCheckpoints);
                    ++decodes;
                }
            }
            __stcg(nxt + i + 1, my_out);
        }
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        if (redo) s_any = 1;
        __syncthreads();
        if (threadIdx.x == 0 && s_any) atomicAdd(&J.ctrl[CTRL_CHANGED + round], 1);
        grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
        const int changed = *reinterpret_cast<volatile int*>(&J.ctrl[CTRL_CHANGED + round]);
        ++round;
        if (changed == 0 || round >= n_sub + kMaxRoundsSlack - 1) break;
    }
    if (decodes) atomicMax(&J.ctrl[CTRL_DECODES], decodes);
    stamp(J, 4);
    // ---- output position of every subsequence: exclusive scan of the blocks each one completes ----
    s_scan[threadIdx.x] = owner ? my_n : 0;
    __syncthreads();
    for (int o = 1; o < kHypThreads; o <<= 1) {
        int t = 0;
        if (threadIdx.x >= o) t = s_scan[threadIdx.x - o];
        __syncthreads();
        s_scan[threadIdx.x] += t;
        __syncthreads();
    }
    if (threadIdx.x == kHypThreads - 1) __stcg(&J.grid_sums[blockIdx.x], s_scan[threadIdx.x]);
    grid_barrier(&J.ctrl[CTRL_BARRIER], gridDim.x);
    long before = 0;
    for (int b = threadIdx.x; b < static_cast<int>(blockIdx.x); b += kHypThreads) before += __ldcg(&J.grid_sums[b]);
    stamp(J, 5);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if ((threadIdx.x & 31) == 0) s_before[threadIdx.x >> 5] = before;
    __syncthreads();
    if (threadIdx.x == 0) {
        long t = 0;
        for (int j = 0; j < kHypThreads / 32; ++j) t += s_before[j];
        s_before[0] = t;
    }
    __syncthreads();
    // write pass: the 8 lanes of a subsequence each take the stretch between two checkpoints of the decode the owner settled on
    const int blk0 = owner ? static_cast<int>(s_before[0] + s_scan[threadIdx.x] - my_n) : 0;
    const int group_leader = static_cast<int>(threadIdx.x & 31u & ~7u);
    const int g_blk0 = __shfl_sync(0xffffffffu, blk0, group_leader);
    const int g_slot = __shfl_sync(0xffffffffu, my_slot, group_leader);
    const uint32_t g_in_x = __shfl_sync(0xffffffffu, my_in.x, group_leader), g_in_y = __shfl_sync(0xffffffffu, my_in.y, group_leader);
    if (active) {
        int n2 = 0;
        if (g_slot >= 0) {
            const uint4* ck = J.ck + (static_cast<size_t>(i) * kLanes + g_slot) * kCheckpoints;
            const uint4 from = __ldcg(ck + lane);
            const uint32_t until = lane + 1 < kCheckpoints ? __ldcg(ck + lane + 1).x : hi;
            decode_range<true>(J, T, make_uint2(from.x, from.y), min(until, hi), n2, static_cast<long>(g_blk0) + from.z);
        } else if (lane == 0) {     // no checkpoints (cannot happen once the loop has settled; kept as the plain path)
            decode_range<true>(J, T, make_uint2(g_in_x, g_in_y), hi, n2, g_blk0);
        }
    }
    if (owner && i == n_sub - 1) {
        const long total = static_cast<long>(blk0) + my_n;
        J.ctrl[CTRL_BLOCKS] = static_cast<int>(total);
        J.ctrl[CTRL_ROUNDS] = round;
        if (total != J.n_blocks) atomicOr(&J.ctrl[CTRL_STATUS], ST_BLOCK_COUNT);
        if (my_out.y != 0) atomicOr(&J.ctrl[CTRL_STATUS], ST_TAIL);
    }
    __syncthreads();
    stamp(J, 6);
}

// ---------------------------------------------------------------------------------------------------------
// DC predictors: exclusive prefix sum of the per-MCU DC differences of one component, reset at every restart interval
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) jpeg_dc_scan_kernel(JpegDev J) {
    __shared__ int sv[1024];
    __shared__ int sf[1024];
    int* dc = J.mcu_dc + blockIdx.x * J.n_mcus;
    const int per = (J.n_mcus + 1023) / 1024;
    const int lo = min(static_cast<int>(threadIdx.x) * per, J.n_mcus), hi = min(lo + per, J.n_mcus);
    int v = 0, f = 0;
    for (int m = lo; m < hi; ++m) {
        if (J.restart && m % J.restart == 0) { v = 0; f = 1; }
        v += dc[m];
    }
    sv[threadIdx.x] = v;
    sf[threadIdx.x] = f;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int lv = 0, lf = 0;
        const bool has = threadIdx.x >= o;
        if (has) { lv = sv[threadIdx.x - o]; lf = sf[threadIdx.x - o]; }
        __syncthreads();
        if (has) {
            if (!sf[threadIdx.x]) sv[threadIdx.x] += lv;
            sf[threadIdx.x] |= lf;
        }
        __syncthreads();
    }
    int run = threadIdx.x ? sv[threadIdx.x - 1] : 0;
    for (int m = lo; m < hi; ++m) {
        if (J.restart && m % J.restart == 0) run = 0;
        const int d = dc[m];
        dc[m] = run;
        run += d;
    }
}

// ---------------------------------------------------------------------------------------------------------
// dequantisation + inverse DCT (jidctint.c: CONST_BITS 13, PASS1_BITS 2), one thread per block
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ unsigned sample_clamp(int v) {
    v = (v + 128) & 1023;              // range_limit[] is indexed modulo 1024 around CENTERJSAMPLE
    return v >= 768 ? 0u : (v > 255 ? 255u : static_cast<unsigned>(v));
}
__device__ __forceinline__ void idct_1d(const int (&in)[8], int (&out)[8], int shift) {
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
    int tmp0 = (in[0] + in[4]) * 8192, tmp1 = (in[0] - in[4]) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7];
    tmp1 = in[5];
    tmp2 = in[3];
    tmp3 = in[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446;
    tmp1 *= 16819;
    tmp2 *= 25172;
    tmp3 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 *= -16069;
    z4 *= -3196;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    out[0] = descale(tmp10 + tmp3, shift);
    out[7] = descale(tmp10 - tmp3, shift);
    out[1] = descale(tmp11 + tmp2, shift);
    out[6] = descale(tmp11 - tmp2, shift);
    out[2] = descale(tmp12 + tmp1, shift);
    out[5] = descale(tmp12 - tmp1, shift);
    out[3] = descale(tmp13 + tmp0, shift);
    out[4] = descale(tmp13 - tmp0, shift);
}

__global__ void __launch_bounds__(128) jpeg_idct_kernel(JpegDev J) {
    __shared__ uint16_t q[3][64];
    for (int i = threadIdx.x; i < 3 * 64; i += 128) q[i >> 6][i & 63] = J.tables->quant[i >> 6][i & 63];
    __syncthreads();
    const long blk = static_cast<long>(blockIdx.x) * 128 + threadIdx.x;
    if (blk >= J.n_blocks) return;
    const int mcu = static_cast<int>(blk / J.bpm), kb = static_cast<int>(blk % J.bpm);
    const int comp = J.tables->comp_of_block[kb], first = J.tables->first_block_of_comp[comp];
    const int j = kb - first;
    const int hs = comp == 0 ? J.hs : 1, vs = comp == 0 ? J.vs : 1;
    const int x = ((mcu % J.mcus_x) * hs + j % hs) * 8, y = ((mcu / J.mcus_x) * vs + j / hs) * 8;
    // the entropy stage stores coefficients at their zig-zag position; the natural order is applied here, at compile time
    constexpr int kNat[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                              41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                              30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    int v[64];
    const uint4* src = reinterpret_cast<const uint4*>(J.coef + blk * 64);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 u = src[r];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int raw = static_cast<int16_t>((w[c >> 1] >> (16 * (c & 1))) & 0xFFFF);
            v[kNat[r * 8 + c]] = raw;
        }
    }
    int dc = J.mcu_dc[comp * J.n_mcus + mcu] + v[0];
    for (int jj = 0; jj < j; ++jj) dc += J.coef[(static_cast<long>(mcu) * J.bpm + first + jj) * 64];
    J.dc_abs[blk] = static_cast<int16_t>(dc);
    v[0] = static_cast<int16_t>(dc);
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] *= q[comp][i];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int in[8], out[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) in[r] = v[r * 8 + c];
        idct_1d(in, out, 13 - 2);
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r * 8 + c] = out[r];
    }
    uint8_t* dst = J.plane[comp] + static_cast<size_t>(y) * J.pitch[comp] + x;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int in[8], out[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) in[c] = v[r * 8 + c];
        idct_1d(in, out, 13 + 2 + 3);
        uint2 px;
        px.x = sample_clamp(out[0]) | (sample_clamp(out[1]) << 8) | (sample_clamp(out[2]) << 16) | (sample_clamp(out[3]) << 24);
        px.y = sample_clamp(out[4]) | (sample_clamp(out[5]) << 8) | (sample_clamp(out[6]) << 16) | (sample_clamp(out[7]) << 24);
        *reinterpret_cast<uint2*>(dst + static_cast<size_t>(r) * J.pitch[comp]) = px;
    }
}

// ---------------------------------------------------------------------------------------------------------
// chroma upsampling + colour conversion -> interleaved BGR, four pixels per thread
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int chroma_at(const JpegDev& J, int comp, int x, int y, int cw, int ch, bool fancy) {
    const uint8_t* p = J.plane[comp];
    const int W = J.pitch[comp];
    if (J.hs == 1) return p[static_cast<size_t>(y) * W + x];
    const int i = x >> 1;
    if (J.vs == 1) {                                           // h2v1
        const uint8_t* row = p + static_cast<size_t>(y) * W;
        if (!fancy) return row[i];
        if (x & 1) return (3 * row[i] + row[min(i + 1, cw - 1)] + 2) >> 2;
        return (3 * row[i] + row[max(i - 1, 0)] + 1) >> 2;
    }
    const int r = y >> 1;                                      // h2v2
    if (!fancy) return p[static_cast<size_t>(r) * W + i];
    const int r1 = (y & 1) ? min(r + 1, ch - 1) : max(r - 1, 0);
    const uint8_t* row0 = p + static_cast<size_t>(r) * W;
    const uint8_t* row1 = p + static_cast<size_t>(r1) * W;
    const int here = 3 * row0[i] + row1[i];
    const int k = (x & 1) ? min(i + 1, cw - 1) : max(i - 1, 0);
    return (3 * here + 3 * row0[k] + row1[k] + ((x & 1) ? 7 : 8)) >> 4;
}

__device__ __forceinline__ unsigned clamp255(int v) { return static_cast<unsigned>(min(max(v, 0), 255)); }

__global__ void __launch_bounds__(256) jpeg_colour_kernel(JpegDev J) {
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4, y = blockIdx.y;
    if (x0 >= J.width) return;
    const int cw = (J.width + J.hs - 1) / J.hs, ch = (J.height + J.vs - 1) / J.vs;
    const bool fancy = cw > 2;
    unsigned px[12];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = min(x0 + j, J.width - 1);
        const int Y = J.plane[0][static_cast<size_t>(y) * J.pitch[0] + x];
        if (J.ncomp == 1) {
            px[3 * j] = px[3 * j + 1] = px[3 * j + 2] = static_cast<unsigned>(Y);
        } else {
            const int cb = chroma_at(J, 1, x, y, cw, ch, fancy) - 128, cr = chroma_at(J, 2, x, y, cw, ch, fancy) - 128;
            px[3 * j + 2] = clamp255(Y + ((91881 * cr + 32768) >> 16));
            px[3 * j + 1] = clamp255(Y + ((-22554 * cb - 46802 * cr + 32768) >> 16));
            px[3 * j + 0] = clamp255(Y + ((116130 * cb + 32768) >> 16));
        }
    }
    uint8_t* dst = J.bgr + static_cast<size_t>(y) * J.stride + static_cast<size_t>(x0) * 3;
    if (x0 + 4 <= J.width && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        d[0] = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
        d[1] = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
        d[2] = px[8] | (px[9] << 8) | (px[10] << 16) | (px[11] << 24);
    } else {
        for (int j = 0; j < 4 && x0 + j < J.width; ++j) {
            dst[3 * j] = static_cast<uint8_t>(px[3 * j]);
            dst[3 * j + 1] = static_cast<uint8_t>(px[3 * j + 1]);
            dst[3 * j + 2] = static_cast<uint8_t>(px[3 * j + 2]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
[[noreturn]] void bad(const std::string& what) { throw std::invalid_argument("jpeg: " + what); }

const uint8_t kNaturalHost[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                  41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                  30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// canonical code tables of one DHT entry -> the kLutBits lookup table and the per-length limits
void build_table(const uint8_t* bits, const uint8_t* vals, bool dc, uint32_t* lut, uint32_t* lut2, uint32_t* base16, int* maxcode,
                 int* valoff, uint8_t* vals_out) {
    std::memset(lut, 0, (1 << kLutBits) * sizeof(uint32_t));
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        valoff[l] = k - code;
        // Kraft bound BEFORE any table write: an over-subscribed length would index past the lookup table
        // (libjpeg rejects such DHT segments, jdhuff.c "Bogus Huffman table definition")
        if (code + bits[l] > (1 << l) || k + bits[l] > 256) bad("over-subscribed Huffman table");
        for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
            if (l <= kLutBits) {
                const int first = code << (kLutBits - l);
                for (int f = 0; f < (1 << (kLutBits - l)); ++f) lut[first + f] = lut_entry(l, vals[k], dc);
            }
        }
        maxcode[l] = bits[l] ? code - 1 : -1;
        if (code > (1 << l)) bad("over-subscribed Huffman table");
        if (l == kLutBits) *base16 = static_cast<uint32_t>(code) << (16 - kLutBits);   // first 16-bit prefix of a longer code
        code <<= 1;
    }
    // second level: every 16-bit prefix from base16 on that starts a code of kLutBits+1 .. 16 bits
    std::memset(lut2, 0, kLut2 * sizeof(uint32_t));
    code = 0;
    k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
            if (l <= kLutBits) continue;
            const uint32_t first = (static_cast<uint32_t>(code) << (16 - l)) - *base16;
            for (uint32_t f = 0; f < (1u << (16 - l)) && first + f < static_cast<uint32_t>(kLut2); ++f) lut2[first + f] = lut_entry(l, vals[k], dc);
        }
        code <<= 1;
    }
    maxcode[0] = -1;
    maxcode[17] = 0x7fffffff;
    valoff[0] = valoff[17] = 0;
    std::memcpy(vals_out, vals, 256);
}

}  // namespace

JpegHeader jpeg_parse_header(const void* file, size_t size) {
    const uint8_t* d = static_cast<const uint8_t*>(file);
    if (d == nullptr || size < 4 || d[0] != 0xFF || d[1] != 0xD8) bad("not a JPEG file (no SOI marker)");
    JpegHeader h;
    bool have_frame = false, have_scan = false;
    size_t i = 2;
    while (!have_scan) {
        if (i + 4 > size) bad("truncated before the scan");
        if (d[i] != 0xFF) bad("marker expected");
        while (i < size && d[i] == 0xFF) ++i;
        if (i >= size) bad("truncated before the scan");
        const int m = d[i++];
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) bad("EOI before any scan");
        if (i + 2 > size) bad("truncated segment");
        const size_t len = (static_cast<size_t>(d[i]) << 8) | d[i + 1];
        if (len < 2 || i + len > size) bad("truncated segment");
        const uint8_t* p = d + i + 2;
        const size_t n = len - 2;
        switch (m) {
            case 0xDB: {
                size_t o = 0;
                while (o < n) {
                    const int prec = p[o] >> 4, id = p[o] & 15;
                    ++o;
                    if (id > 3 || o + (prec ? 128u : 64u) > n) bad("bad DQT segment");
                    for (int k = 0; k < 64; ++k) {
                        h.quant[id][kNaturalHost[k]] = prec ? static_cast<uint16_t>((p[o] << 8) | p[o + 1]) : p[o];
                        o += prec ? 2 : 1;
                    }
                    h.have_quant[id] = true;
                }
                break;
            }
            case 0xC4: {
                size_t o = 0;
                while (o < n) {
                    if (o + 17 > n) bad("bad DHT segment");
                    const int tc = p[o] >> 4, id = p[o] & 15;
                    if (tc > 1 || id > 3) bad("bad DHT table id");
                    int cnt = 0;
                    h.bits[tc][id][0] = 0;
                    for (int l = 1; l <= 16; ++l) cnt += (h.bits[tc][id][l] = p[o + l]);
                    o += 17;
                    if (cnt > 256 || o + cnt > n) bad("bad DHT segment");
                    std::memset(h.vals[tc][id], 0, 256);
                    std::memcpy(h.vals[tc][id], p + o, cnt);
                    o += cnt;
                    h.have_huff[tc][id] = true;
                }
                break;
            }
            case 0xC0:
            case 0xC1: {
                if (n < 6) bad("bad SOF segment");
                if (p[0] != 8) bad("only 8-bit samples are supported");
                h.height = (p[1] << 8) | p[2];
                h.width = (p[3] << 8) | p[4];
                h.components = p[5];
                if (h.components != 1 && h.components != 3) bad("only 1 or 3 components are supported");
                if (n < static_cast<size_t>(6 + 3 * h.components) || h.width == 0 || h.height == 0) bad("bad SOF segment");
                int hs[3] = {1, 1, 1}, vs[3] = {1, 1, 1};
                for (int c = 0; c < h.components; ++c) {
                    h.component_id[c] = p[6 + 3 * c];
                    hs[c] = p[7 + 3 * c] >> 4;
                    vs[c] = p[7 + 3 * c] & 15;
                    h.quant_of[c] = p[8 + 3 * c];
                    if (h.quant_of[c] > 3) bad("bad quantisation table id");
                }
                if (h.components == 1) {
                    h.h_samp = h.v_samp = 1;   // a one-component scan is never interleaved
                } else {
                    if (hs[1] != 1 || vs[1] != 1 || hs[2] != 1 || vs[2] != 1) bad("unsupported chroma sampling");
                    if (!((hs[0] == 1 && vs[0] == 1) || (hs[0] == 2 && vs[0] == 1) || (hs[0] == 2 && vs[0] == 2)))
                        bad("unsupported luma sampling (4:4:4, 4:2:2 and 4:2:0 are supported)");
                    h.h_samp = hs[0];
                    h.v_samp = vs[0];
                }
                have_frame = true;
                break;
            }
            case 0xDD:
                if (n < 2) bad("bad DRI segment");
                h.restart_interval = (p[0] << 8) | p[1];
                break;
            case 0xDA: {
                if (!have_frame) bad("SOS before SOF");
                if (n < 1 || p[0] != h.components || n < static_cast<size_t>(4 + 2 * h.components))
                    bad("only one interleaved scan over all components is supported");
                for (int c = 0; c < h.components; ++c) {
                    h.dc_of[c] = p[2 + 2 * c] >> 4;
                    h.ac_of[c] = p[2 + 2 * c] & 15;
                    if (h.dc_of[c] > 3 || h.ac_of[c] > 3) bad("bad Huffman table id");
                }
                h.scan_offset = i + len;
                have_scan = true;
                break;
            }
            case 0xE1: {   // APP1: cv::imread applies the EXIF orientation (it rotates / mirrors the frame); only "normal" is accepted
                if (n >= 14 && std::memcmp(p, "Exif\0\0", 6) == 0) {
                    const uint8_t* t = p + 6;
                    const size_t tn = n - 6;
                    const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
                    auto u16 = [&](size_t o) { return le ? t[o] | (t[o + 1] << 8) : (t[o] << 8) | t[o + 1]; };
                    auto u32 = [&](size_t o) {
                        return le ? static_cast<size_t>(t[o]) | (static_cast<size_t>(t[o + 1]) << 8) | (static_cast<size_t>(t[o + 2]) << 16) |
                                        (static_cast<size_t>(t[o + 3]) << 24)
                                  : (static_cast<size_t>(t[o]) << 24) | (static_cast<size_t>(t[o + 1]) << 16) |
                                        (static_cast<size_t>(t[o + 2]) << 8) | static_cast<size_t>(t[o + 3]);
                    };
                    if ((le || be) && u16(2) == 42) {
                        const size_t ifd = u32(4);
                        if (ifd + 2 <= tn) {
                            const int entries = u16(ifd);
                            for (int e = 0; e < entries && ifd + 2 + static_cast<size_t>(e + 1) * 12 <= tn; ++e) {
                                const size_t o = ifd + 2 + static_cast<size_t>(e) * 12;
                                if (u16(o) == 0x0112) {
                                    const int orientation = u16(o + 8);
                                    if (orientation > 1 && orientation <= 8)
                                        bad("EXIF orientation " + std::to_string(orientation) +
                                            " is not supported (cv::imread would rotate or mirror this frame)");
                                }
                            }
                        }
                    }
                }
                break;
            }
            case 0xE0:     // APP0 "JFIF": three components are YCbCr by definition
                if (n >= 5 && std::memcmp(p, "JFIF", 5) == 0) h.jfif = true;
                break;
            case 0xEE:     // APP14 "Adobe": without JFIF, transform 0 on three components means the samples are RGB
                if (n >= 12 && std::memcmp(p, "Adobe", 5) == 0) h.adobe_transform = p[11];
                break;
            default:
                if (m >= 0xC2 && m <= 0xCF && m != 0xC8 && m != 0xCC)
                    bad("only baseline / extended sequential Huffman JPEG is supported (this file is progressive, lossless or arithmetic)");
                break;   // APPn, COM, ...
        }
        i += len;
    }
    for (int c = 0; c < h.components; ++c)
        if (!h.have_quant[h.quant_of[c]] || !h.have_huff[0][h.dc_of[c]] || !h.have_huff[1][h.ac_of[c]]) bad("missing table");
    // libjpeg's colour-space guess (jdapimin.c default_decompress_parms): JFIF means YCbCr; otherwise an Adobe marker with
    // transform 0 means RGB; with neither marker, component ids that spell "RGB" mean RGB
    if (h.components == 3 && !h.jfif &&
        (h.adobe_transform == 0 ||
         (h.adobe_transform < 0 && h.component_id[0] == 'R' && h.component_id[1] == 'G' && h.component_id[2] == 'B')))
        bad("RGB-coded JPEG (no YCbCr transform) is not supported");
    // the entropy-coded segment ends at EOI: the last FF D9 of the file
    size_t end = size;
    while (end >= h.scan_offset + 2 && !(d[end - 2] == 0xFF && d[end - 1] == 0xD9)) --end;
    if (end < h.scan_offset + 2) bad("no EOI marker");
    h.scan_bytes = end - 2 - h.scan_offset;
    if (h.scan_bytes == 0) bad("empty scan");
    if (h.scan_bytes >= (1ull << 28)) bad("scan larger than 256 MB");
    const int mcu_w = 8 * h.h_samp, mcu_h = 8 * h.v_samp;
    h.mcus_x = (h.width + mcu_w - 1) / mcu_w;
    h.mcus_y = (h.height + mcu_h - 1) / mcu_h;
    h.blocks_per_mcu = h.h_samp * h.v_samp + (h.components == 3 ? 2 : 0);
    h.n_mcus = static_cast<long>(h.mcus_x) * h.mcus_y;
    h.n_blocks = h.n_mcus * h.blocks_per_mcu;
    h.n_intervals = h.restart_interval ? static_cast<int>((h.n_mcus + h.restart_interval - 1) / h.restart_interval) : 1;
    return h;
}

namespace {
struct PinnedBlock {   // host staging: the tables followed by the scan bytes
    JpegTables tables;
};
}  // namespace

JpegDecoder::JpegDecoder(int device) : device_(device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        throw std::runtime_error("rm_radar_b200: no CUDA device (sm_100 required; there is no CPU fallback)");
    if (device < 0 || device >= n) throw std::invalid_argument("JpegDecoder: bad device index");
    RMR_CUDA(cudaSetDevice(device_));
    cudaDeviceProp prop{};
    RMR_CUDA(cudaGetDeviceProperties(&prop, device_));
    if (prop.major != 10) throw std::runtime_error("rm_radar_b200: sm_100 (B200) required, found sm_" + std::to_string(prop.major * 10 + prop.minor));
    int per_sm = 0;
    RMR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jpeg_entropy_simple_kernel, kDecodeThreads, 0));
    max_coresident_ = per_sm * prop.multiProcessorCount;
    RMR_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    RMR_CUDA(cudaEventCreateWithFlags(&staged_, cudaEventDisableTiming));
    RMR_CUDA(cudaMalloc(&tables_, sizeof(JpegTables)));
    RMR_CUDA(cudaMalloc(&bmap_, kMaxHypBlocks * kLanes));
    RMR_CUDA(cudaMalloc(&tstamp_, 8 * sizeof(unsigned long long)));
    RMR_CUDA(cudaMemset(tstamp_, 0, 8 * sizeof(unsigned long long)));
    {
        const char* e = std::getenv("RMR_JPEG_SIMPLE");
        simple_ = e && e[0] == '1';
        int hyp_per_sm = 0;
        RMR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&hyp_per_sm, jpeg_entropy_kernel, kHypThreads, 0));
        max_hyp_blocks_ = std::min(hyp_per_sm * prop.multiProcessorCount, kMaxHypBlocks);
    }
    RMR_CUDA(cudaMallocHost(&pinned_status_, 8 * sizeof(int)));
    std::memset(pinned_status_, 0, 8 * sizeof(int));
}

JpegDecoder::~JpegDecoder() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    cudaFreeHost(pinned_raw_);
    cudaFree(dev_raw_);
    cudaFree(dev_stream_);
    cudaFree(blk_counts_);
    cudaFree(blk_offsets_);
    cudaFree(intervals_);
    cudaFree(states_);
    cudaFree(cand_);
    cudaFree(res_);
    cudaFree(ck_);
    cudaFree(bmap_);
    cudaFree(tstamp_);
    cudaFree(changed_);
    cudaFree(coef_);
    cudaFree(dc_abs_);
    cudaFree(mcu_dc_);
    cudaFree(planes_);
    cudaFree(frame_);
    cudaFree(tables_);
    cudaFreeHost(pinned_status_);
    if (staged_) cudaEventDestroy(staged_);
    for (auto& e : stage_ev_)
        if (e) cudaEventDestroy(e);
    // the stream is owned only while nobody replaced it; destroying a foreign stream is the caller's business
}

namespace {
template <typename T>
void regrow(T*& p, size_t count, bool pinned = false) {
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; }
    if (pinned) RMR_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p), count * sizeof(T)));
    else RMR_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
}
inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

void JpegDecoder::reserve(const JpegHeader& h) {
    const size_t raw = round_up(h.scan_bytes + 64, 4096);
    if (raw > cap_raw_) {
        RMR_CUDA(cudaStreamSynchronize(stream_));
        cap_raw_ = raw + raw / 4;
        regrow(pinned_raw_, sizeof(JpegTables) + cap_raw_, true);
        regrow(dev_raw_, cap_raw_);
        regrow(dev_stream_, cap_raw_);
        const size_t ub = cap_raw_ / (kUnstuffThreads * kUnstuffBytes) + 1;
        regrow(blk_counts_, ub);
        regrow(blk_offsets_, ub);
        cap_sub_ = static_cast<long>(cap_raw_ * 8 / 256 + 2);      // never cut finer than 256 bits
        regrow(states_, 2 * static_cast<size_t>(cap_sub_ + 1));
        regrow(cand_, static_cast<size_t>(cap_sub_ + 1) * kLanes);
        regrow(res_, static_cast<size_t>(cap_sub_) * kLanes);
        regrow(ck_, static_cast<size_t>(cap_sub_) * kLanes * kCheckpoints);
        regrow(changed_, static_cast<size_t>(CTRL_CHANGED + cap_sub_ + kMaxRoundsSlack + max_coresident_ + 64));
    }
    if (h.n_blocks > cap_blocks_) {
        RMR_CUDA(cudaStreamSynchronize(stream_));
        cap_blocks_ = h.n_blocks;
        regrow(coef_, static_cast<size_t>(cap_blocks_) * 64);
        regrow(dc_abs_, static_cast<size_t>(cap_blocks_));
        regrow(planes_, static_cast<size_t>(cap_blocks_) * 64);
    }
    if (h.n_mcus > cap_mcus_) {
        RMR_CUDA(cudaStreamSynchronize(stream_));
        cap_mcus_ = h.n_mcus;
        regrow(mcu_dc_, static_cast<size_t>(cap_mcus_) * 3);
    }
    if (h.n_intervals + 1 > cap_intervals_) {
        RMR_CUDA(cudaStreamSynchronize(stream_));
        cap_intervals_ = h.n_intervals + 1;
        regrow(intervals_, static_cast<size_t>(cap_intervals_) + 1);
    }
}

const uint8_t* JpegDecoder::decode(const void* file, size_t size, uint8_t* dev_bgr, int stride, int* width, int* height) {
    RMR_CUDA(cudaSetDevice(device_));
    const JpegHeader h = jpeg_parse_header(file, size);
    reserve(h);
    if (dev_bgr == nullptr) {
        const size_t need = static_cast<size_t>(h.width) * h.height * 3;
        if (need > cap_frame_) {
            RMR_CUDA(cudaStreamSynchronize(stream_));
            cap_frame_ = need;
            regrow(frame_, need + 16);
        }
        dev_bgr = frame_;
        stride = h.width * 3;
    } else if (stride < h.width * 3) {
        throw std::invalid_argument("jpeg: output stride smaller than width * 3");
    }
    last_ = h;

    // ---- staging: tables + scan bytes -> pinned -> device (one block each) ----
    RMR_CUDA(cudaEventSynchronize(staged_));
    int stage = 0;
    auto mark = [&] { if (profiling_) RMR_CUDA(cudaEventRecord(stage_ev_[stage++], stream_)); };
    mark();
    JpegTables* T = reinterpret_cast<JpegTables*>(pinned_raw_);
    std::memset(T, 0, sizeof(JpegTables));
    int kb = 0;
    for (int c = 0; c < h.components; ++c) {
        build_table(h.bits[0][h.dc_of[c]], h.vals[0][h.dc_of[c]], true, T->lut[2 * c], T->lut2[2 * c], &T->base16[2 * c], T->maxcode[2 * c],
                    T->valoff[2 * c], T->vals[2 * c]);
        build_table(h.bits[1][h.ac_of[c]], h.vals[1][h.ac_of[c]], false, T->lut[2 * c + 1], T->lut2[2 * c + 1], &T->base16[2 * c + 1],
                    T->maxcode[2 * c + 1], T->valoff[2 * c + 1], T->vals[2 * c + 1]);
        std::memcpy(T->quant[c], h.quant[h.quant_of[c]], sizeof(T->quant[c]));
        T->first_block_of_comp[c] = static_cast<uint8_t>(kb);
        const int nb = c == 0 ? h.h_samp * h.v_samp : 1;
        for (int j = 0; j < nb; ++j) T->comp_of_block[kb++] = static_cast<uint8_t>(c);
    }
    uint8_t* raw_pinned = pinned_raw_ + sizeof(JpegTables);
    std::memcpy(raw_pinned, static_cast<const uint8_t*>(file) + h.scan_offset, h.scan_bytes);
    std::memset(raw_pinned + h.scan_bytes, 0, 32);
    RMR_CUDA(cudaMemcpyAsync(tables_, T, sizeof(JpegTables), cudaMemcpyHostToDevice, stream_));
    RMR_CUDA(cudaMemcpyAsync(dev_raw_, raw_pinned, h.scan_bytes + 32, cudaMemcpyHostToDevice, stream_));
    RMR_CUDA(cudaEventRecord(staged_, stream_));
    last_upload_ = sizeof(JpegTables) + h.scan_bytes + 32;

    // ---- geometry ----
    JpegDev J{};
    J.raw = dev_raw_;
    J.n_raw = static_cast<uint32_t>(h.scan_bytes);
    J.stream = dev_stream_;
    J.blk_counts = blk_counts_;
    J.blk_offsets = blk_offsets_;
    J.n_ublocks = static_cast<int>((h.scan_bytes + kUnstuffThreads * kUnstuffBytes - 1) / (kUnstuffThreads * kUnstuffBytes));
    J.intervals = intervals_;
    J.n_intervals = h.n_intervals;
    J.states = states_;
    static const int env_bits = [] { const char* e = std::getenv("RMR_JPEG_SUB_BITS"); return e ? std::atoi(e) : 0; }();
    uint32_t sub_bits = env_bits >= 256 ? static_cast<uint32_t>(env_bits) / 32 * 32 : 1024;
    const unsigned long long raw_bits = static_cast<unsigned long long>(h.scan_bytes) * 8;
    // every subsequence needs a resident thread (simple kernel) or 8 resident lanes (hypothesis kernel)
    const unsigned long long max_subs = simple_ ? static_cast<unsigned long long>(max_coresident_) * kDecodeThreads
                                                : static_cast<unsigned long long>(max_hyp_blocks_) * kSubsPerBlock;
    if ((raw_bits + sub_bits - 1) / sub_bits > max_subs)
        sub_bits = static_cast<uint32_t>(round_up((raw_bits + max_subs - 1) / max_subs, 32));
    J.sub_bits = sub_bits;
    J.n_sub = static_cast<int>((raw_bits + sub_bits - 1) / sub_bits);
    J.cand = cand_;
    J.res = res_;
    J.bmap = bmap_;
    J.ck = ck_;
    J.tstamp = tstamp_;
    J.ctrl = changed_;
    J.grid_sums = changed_ + CTRL_CHANGED + cap_sub_ + kMaxRoundsSlack;
    J.coef = coef_;
    J.dc_abs = dc_abs_;
    J.mcu_dc = mcu_dc_;
    J.n_blocks = h.n_blocks;
    J.n_mcus = static_cast<int>(h.n_mcus);
    J.bpm = h.blocks_per_mcu;
    J.ncomp = h.components;
    J.restart = h.restart_interval;
    J.tables = static_cast<const JpegTables*>(tables_);
    J.width = h.width;
    J.height = h.height;
    J.mcus_x = h.mcus_x;
    J.mcus_y = h.mcus_y;
    J.hs = h.h_samp;
    J.vs = h.v_samp;
    size_t off = 0;
    for (int c = 0; c < h.components; ++c) {
        const int hs = c == 0 ? h.h_samp : 1, vs = c == 0 ? h.v_samp : 1;
        J.pitch[c] = h.mcus_x * 8 * hs;
        J.plane[c] = planes_ + off;
        off += static_cast<size_t>(J.pitch[c]) * h.mcus_y * 8 * vs;
    }
    J.bgr = dev_bgr;
    J.stride = stride;

    // ---- launches ----
    mark();
    RMR_CUDA(cudaMemsetAsync(changed_, 0, sizeof(int) * static_cast<size_t>(CTRL_CHANGED + J.n_sub + kMaxRoundsSlack), stream_));
    RMR_CUDA(cudaMemsetAsync(coef_, 0, static_cast<size_t>(h.n_blocks) * 64 * sizeof(int16_t), stream_));
    RMR_CUDA(cudaMemsetAsync(mcu_dc_, 0, static_cast<size_t>(h.n_mcus) * h.components * sizeof(int), stream_));
    mark();
    jpeg_unstuff_count_kernel<<<J.n_ublocks, kUnstuffThreads, 0, stream_>>>(J);
    jpeg_unstuff_scan_kernel<<<1, 1024, 0, stream_>>>(J);
    jpeg_unstuff_scatter_kernel<<<J.n_ublocks, kUnstuffThreads, 0, stream_>>>(J);
    mark();
    {
        void* args[] = {&J};
        if (simple_) {
            const int grid = (J.n_sub + kDecodeThreads - 1) / kDecodeThreads;
            RMR_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(jpeg_entropy_simple_kernel), dim3(grid),
                                                 dim3(kDecodeThreads), args, 0, stream_));
        } else {
            const int grid = (J.n_sub + kSubsPerBlock - 1) / kSubsPerBlock;
            RMR_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(jpeg_entropy_kernel), dim3(grid), dim3(kHypThreads),
                                                 args, 0, stream_));
        }
    }
    mark();
    jpeg_dc_scan_kernel<<<h.components, 1024, 0, stream_>>>(J);
    mark();
    jpeg_idct_kernel<<<static_cast<int>((h.n_blocks + 127) / 128), 128, 0, stream_>>>(J);
    mark();
    jpeg_colour_kernel<<<dim3((h.width + 1023) / 1024, h.height), 256, 0, stream_>>>(J);
    mark();
    RMR_CUDA(cudaGetLastError());
    RMR_CUDA(cudaMemcpyAsync(pinned_status_, changed_, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    last_launches_ = 7;
    if (width) *width = h.width;
    if (height) *height = h.height;
    return dev_bgr;
}

void JpegDecoder::profile(const void* file, size_t size, float* stage_ms) {
    for (auto& e : stage_ev_)
        if (!e) RMR_CUDA(cudaEventCreate(&e));
    profiling_ = true;
    try {
        decode(file, size, nullptr, 0, nullptr, nullptr);
    } catch (...) {
        profiling_ = false;
        throw;
    }
    profiling_ = false;
    RMR_CUDA(cudaStreamSynchronize(stream_));
    for (int i = 0; i < kStages; ++i) RMR_CUDA(cudaEventElapsedTime(&stage_ms[i], stage_ev_[i], stage_ev_[i + 1]));
    // phases of the entropy kernel as block 0 saw them: pass 0, pass 1 + chase level 1, chase + seed, verify loop, scan, write
    unsigned long long t[8] = {};
    RMR_CUDA(cudaMemcpy(t, tstamp_, sizeof(t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 6; ++i) stage_ms[kStages + i] = simple_ ? 0.f : static_cast<float>(static_cast<double>(t[i + 1] - t[i]) * 1e-6);
}

int JpegDecoder::status() {
    RMR_CUDA(cudaStreamSynchronize(stream_));
    return pinned_status_[CTRL_STATUS];
}

int JpegDecoder::last_loop_decodes() {
    RMR_CUDA(cudaStreamSynchronize(stream_));
    return pinned_status_[CTRL_DECODES];
}

int JpegDecoder::last_rounds() {
    RMR_CUDA(cudaStreamSynchronize(stream_));
    return pinned_status_[CTRL_ROUNDS];
}

void JpegDecoder::decode_to_host(const void* file, size_t size, uint8_t* host_bgr, size_t capacity, int* width, int* height) {
    int w = 0, h = 0;
    const uint8_t* dev = decode(file, size, nullptr, 0, &w, &h);
    const size_t need = static_cast<size_t>(w) * h * 3;
    if (width) *width = w;
    if (height) *height = h;
    if (host_bgr == nullptr || capacity < need) throw std::invalid_argument("jpeg: host buffer smaller than width * height * 3");
    RMR_CUDA(cudaMemcpyAsync(host_bgr, dev, need, cudaMemcpyDeviceToHost, stream_));
    const int st = status();
    if (st) throw std::runtime_error("jpeg: corrupt entropy-coded data (status " + std::to_string(st) + ")");
}

long JpegDecoder::read_coefficients(int16_t* out, long capacity_blocks) {
    const long n = last_.n_blocks;
    if (out == nullptr || capacity_blocks < n) throw std::invalid_argument("jpeg: coefficient buffer too small");
    RMR_CUDA(cudaStreamSynchronize(stream_));
    std::vector<int16_t> zz(static_cast<size_t>(n) * 64), dc(static_cast<size_t>(n));
    RMR_CUDA(cudaMemcpy(zz.data(), coef_, zz.size() * sizeof(int16_t), cudaMemcpyDeviceToHost));
    RMR_CUDA(cudaMemcpy(dc.data(), dc_abs_, dc.size() * sizeof(int16_t), cudaMemcpyDeviceToHost));
    for (long b = 0; b < n; ++b) {        // device layout: zig-zag order, DC as a difference
        for (int k = 0; k < 64; ++k) out[b * 64 + kNaturalHost[k]] = zz[static_cast<size_t>(b) * 64 + k];
        out[b * 64] = dc[static_cast<size_t>(b)];
    }
    return n;
}

}  // namespace rmr
