// Convolution layer launchers (implicit GEMM on tcgen05 + SIMT stem / checker + pooling ops).
#pragma once
#include "common.cuh"

namespace rmr {

// One Conv(+bias+SiLU+residual) op at a fixed batch size; activations NHWC fp16 with channel pitch.
struct ConvDesc {
    const __half* in = nullptr;
    int in_pitch = 0, in_coff = 0, cin = 0, h_in = 0, w_in = 0;
    void* out = nullptr;
    int out_pitch = 0, out_coff = 0, cout = 0, out_f32 = 0, h_out = 0, w_out = 0;
    int k = 1, stride = 1, act = 0;
    const __half* res = nullptr;
    int res_pitch = 0, res_coff = 0;
    // optional second destination of the same fp16 result: a plain copy (dup_mode 1: Concat input with two
    // homes) or a nearest-neighbour 2x upsample (dup_mode 2: each pixel lands on its 2x2 block)
    __half* dup = nullptr;
    int dup_pitch = 0, dup_coff = 0, dup_mode = 0;
    const __half* w = nullptr;   // [cout_pad][k*k][cin_pad] fp16
    const float* bias = nullptr; // [cout_pad]
    int cout_pad = 0, cin_pad = 0;
    int n = 1;                   // batch
};

struct ConvParams {
    int n, h_out, w_out, cout;
    int block_n, bk, kpt, ntaps;      // kpt = k-blocks per tap
    int tw, th, tn, tiles_w, tiles_h, tiles_n;
    int4 tap[9];                      // (channel add, dw, h-parity, dh)
    int cin_coff, cin;
    void* out;
    int out_pitch, out_coff, out_f32;
    const float* bias;
    int act;
    const __half* res;
    int res_pitch, res_coff;
    __half* dup;
    int dup_pitch, dup_coff, dup_mode;
    uint32_t idesc, sbo, layout;
    uint32_t a_bytes, b_bytes, b_off, stage_stride, tmem_cols;   // operand ring geometry
    int stages, vec_ok;
    int dual;                            // MMA issue streams: 0/1 = one (warp 4), 2 = +warp 6, 4 = +warps 7, 8; accumulators acc_stride columns apart
    uint32_t acc_stride;
    int slim;                            // 192-thread / 3-CTAs-per-SM kernel variant
    int pair;                            // CTA-pair mode (cta_group::2): b_bytes / tm_b box hold half of the weight rows
    int halo, na;                        // halo mode (3x3 stride 1): on/off, activation patch slots (1 or 2)
    int splits, it_per_split, part_ld;   // split-K: grid.z splits of it_per_split k-blocks; partial row pitch
    float* partial;                      // [tile][split][128][part_ld] fp32 partial tiles
    int* counters;                       // [tile] arrivals, zero between launches
    int dbg_flags;                    // profiling only: bit0 = issue no MMAs, bit1 = one k-step per k-block
    long long* dbg;                   // optional per-CTA clock64 timeline (64 slots per CTA), tests only
};

// Round-2 kernel (conv2.cu): persistent CTAs, resident weights, halo patches, double-buffered TMEM.
struct Conv2Params {
    int n, h_out, w_out, cout;
    int tw, th, tn, tiles_w, tiles_h, tiles_n, m_tiles;
    int block_n, n_tiles, splits, ns_total;   // ns_total = n_tiles * splits: (channel slice, K split) pairs
    int gm;                                   // CTAs per pair; a CTA takes pixel tiles j, j + gm, ...
    int in_pitch;                             // channel pitch of the input buffer (stride-2 halo: column parity = a channel offset)
    uint32_t cls_stride;                      // stride-2 halo: bytes between the four parity-class patches of a stage
    int exit_wait_all;                        // 1: the DMA warp waits for its bulk stores to complete before the CTA retires
    int gm_w, gm_h, gm_n;                     // gm as a step in (tile_w, tile_h, tile_n): the kernel walks tiles without dividing
    int units_per_split;                      // split granularity: (tap, chunk) pairs, or channel chunks in halo mode
    int bk, kpt, ntaps, cin, cin_coff;
    int halo, pw;                             // halo mode: one [18][pw]-pixel patch per channel chunk
    uint32_t row_bytes;                       // bk * 2
    int4 tap[9];                              // (channel add, dw, h-parity, dh)
    uint32_t a_stage, a_tx, b_stage, a_off, b_off;   // shared-memory geometry (bytes): stage strides, ring offsets
    uint32_t a_kb, b_kb, b_in_stage;                 // bytes of one k-block of A / B; joint ring: offset of the weights inside a stage
    int sa, sb, b_resident;                          // A (or joint) stages; B stages (halo streaming: three taps each)
    int g, joint;                                    // per-tap mode: k-blocks per stage; weights travel in the A stage
    uint32_t idesc, sbo_a, sbo_b, layout;
    uint32_t acc_stride, tmem_cols;
    void* out;
    int out_pitch, out_coff, out_f32;
    const float* bias;
    int act;                                  // 0 none, 1 SiLU (tanh form), 2 SiLU (ex2 + rcp)
    const __half* res;
    int res_pitch, res_coff;
    __half* dup;
    int dup_pitch, dup_coff, dup_mode;
    int vec_ok;
    int part_ld;
    float* partial;
    int* counters;
    long long* dbg;                           // optional per-CTA clock64 timeline (64 slots per CTA), tests only
    int dbg_mode;                             // 0: per-tile phases; 1: per-k-block stamps of the first tile (A issue 4.., B issue 24.., consumed 44..)
    // epilogue through shared memory + TMA (one [128 rows][32 channels] sub-tile per 32-column chunk)
    int tma_epi, nchunks, esize_out;
    uint32_t stage_off, sub_bytes, chunk_bytes;
};

struct ConvLaunch {
    CUtensorMap tm_a, tm_b;
    CUtensorMap tm_out, tm_res;           // conv2 TMA epilogue: output tile store, shortcut tile load
    CUtensorMap tm_dup[4];                // second destination: [0] plain copy, or the four (dy, dx) phases of the 2x upsample
    ConvParams p;
    Conv2Params q;
    int v2 = 0;          // 1: launch conv2_kernel with q; 0: the round-1 kernel with p
    dim3 grid;
    int smem_bytes = 0;
    double flops = 0;
};

void conv_init();   // one-time function attributes; must run outside stream capture
bool conv_umma_supported(const ConvDesc& d);
ConvLaunch make_conv_launch(const ConvDesc& d);
// scratch a split-K launch needs (0 when the layer is not split); bind a zero-initialised region before launching
size_t conv_scratch_bytes(const ConvLaunch& l);
void conv_bind_scratch(ConvLaunch& l, void* zeroed_base);
void launch_conv_umma(const ConvLaunch& l, cudaStream_t s, bool pdl = true);
// conv2.cu
bool conv2_enabled();
bool conv2_supported(const ConvDesc& d);
void plan_conv2(const ConvDesc& d, ConvLaunch& l);      // host-only planning (no CUDA calls)
void make_conv2_launch(const ConvDesc& d, ConvLaunch& l);
size_t conv2_scratch_bytes(const ConvLaunch& l);
void conv2_bind_scratch(ConvLaunch& l, void* zeroed_base);
void conv2_init();
void launch_conv2(const ConvLaunch& l, cudaStream_t s, bool pdl);
// generic direct convolution on CUDA cores: the stem (Cin=3) and the on-device checker for tests
void launch_conv_simt(const ConvDesc& d, cudaStream_t s);

void launch_maxpool5(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff,
                     int n, int h, int w, int c, cudaStream_t s);
// SPPF's chained 5x5 max-pools (y1, y2, y3 -> three channel offsets of one buffer) in one launch; h*w <= 1024
void launch_sppf_pool3(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int coff1, int coff2,
                       int coff3, int n, int h, int w, int c, cudaStream_t s);
void launch_upsample2(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff,
                      int n, int h_in, int w_in, int c, cudaStream_t s);
void launch_copy_channels(const __half* in, int in_pitch, int in_coff, __half* out, int out_pitch, int out_coff,
                          int n, int h, int w, int c, cudaStream_t s);

}  // namespace rmr
