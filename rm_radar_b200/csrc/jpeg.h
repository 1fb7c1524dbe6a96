// Baseline JPEG file image -> BGR8 frame in device memory (see jpeg.cu).
#pragma once
#include <cstddef>
#include <cstdint>

#include "common.cuh"

namespace rmr {

struct JpegHeader {
    int width = 0, height = 0, components = 0;
    int h_samp = 1, v_samp = 1;          // luma sampling factors (chroma is 1x1)
    int restart_interval = 0;            // MCUs per restart interval, 0 = none
    int component_id[3] = {0, 0, 0};
    int adobe_transform = -1;            // APP14 transform byte, -1 = no Adobe marker
    bool jfif = false;                   // APP0 JFIF marker seen
    int quant_of[3] = {0, 0, 0}, dc_of[3] = {0, 0, 0}, ac_of[3] = {0, 0, 0};
    uint16_t quant[4][64] = {};          // natural (row-major) order
    uint8_t bits[2][4][17] = {};         // [dc/ac][table id][code length] counts
    uint8_t vals[2][4][256] = {};
    bool have_quant[4] = {}, have_huff[2][4] = {};
    size_t scan_offset = 0, scan_bytes = 0;   // entropy-coded segment: [scan_offset, scan_offset + scan_bytes) ends before EOI
    // derived
    int mcus_x = 0, mcus_y = 0, blocks_per_mcu = 0;
    long n_mcus = 0, n_blocks = 0;
    int n_intervals = 1;
};

// host: marker segments (SOI .. SOS).  Throws std::invalid_argument for anything that is not a baseline /
// extended-sequential 8-bit Huffman file with one interleaved scan and 4:4:4 / 4:2:2 / 4:2:0 / grayscale sampling.
JpegHeader jpeg_parse_header(const void* file, size_t size);

class JpegDecoder {
public:
    explicit JpegDecoder(int device);
    ~JpegDecoder();
    JpegDecoder(const JpegDecoder&) = delete;
    JpegDecoder& operator=(const JpegDecoder&) = delete;

    void set_stream(cudaStream_t s) { stream_ = s; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }

    // Enqueues upload + decode of one file on the decoder's stream.  The frame is written to `dev_bgr` (row pitch
    // `stride` bytes) or, when dev_bgr is null, to the decoder's own frame buffer (pitch width * 3).  Returns the
    // device pointer of the frame; nothing is synchronised.
    const uint8_t* decode(const void* file, size_t size, uint8_t* dev_bgr, int stride, int* width, int* height);
    // waits for the last decode and reports whether the stream decoded cleanly (0) or which check failed
    int status();
    // cv::imread equivalent: decode + copy to host + status check (throws std::runtime_error on a corrupt stream)
    void decode_to_host(const void* file, size_t size, uint8_t* host_bgr, size_t capacity, int* width, int* height);
    // test hook: quantised coefficient blocks of the last decode, scan order x natural order, DC resolved
    long read_coefficients(int16_t* out, long capacity_blocks);
    int last_launches() const { return last_launches_; }
    int last_rounds();                  // synchronisation rounds the entropy decoder needed (after status())
    int last_loop_decodes();            // most re-decodes any one thread did in the verification loop
    size_t last_upload_bytes() const { return last_upload_; }
    // stage timing of one decode (CUDA events between the launches): upload, clear, unstuff, entropy, dc scan, idct, colour;
    // stage_ms holds kStages + 6 floats: the last six split the entropy kernel (globaltimer stamps of block 0)
    static constexpr int kStages = 7;
    void profile(const void* file, size_t size, float* stage_ms);

private:
    void reserve(const JpegHeader& h);
    int device_;
    cudaStream_t stream_ = nullptr;
    JpegHeader last_{};
    // sizes the buffers were allocated for
    size_t cap_raw_ = 0;
    long cap_blocks_ = 0, cap_mcus_ = 0, cap_sub_ = 0;
    int cap_intervals_ = 0;
    size_t cap_frame_ = 0, cap_planes_ = 0;
    uint8_t *pinned_raw_ = nullptr, *dev_raw_ = nullptr, *dev_stream_ = nullptr;
    int2 *blk_counts_ = nullptr, *blk_offsets_ = nullptr;
    uint32_t* intervals_ = nullptr;     // start bit of every restart interval + the end of the stream
    uint2* states_ = nullptr;           // [2][n_sub + 1]
    uint2* cand_ = nullptr;             // [n_sub + 1][8] candidate states per subsequence boundary
    uint4* res_ = nullptr;              // [n_sub][8] decode results from each candidate
    uint4* ck_ = nullptr;               // [n_sub][8][8] checkpoints inside each of those decodes (write pass split)
    unsigned char* bmap_ = nullptr;     // per-block lane maps of the link chase
    unsigned long long* tstamp_ = nullptr;
    bool simple_ = false;               // RMR_JPEG_SIMPLE=1: single-hypothesis kernel
    int max_hyp_blocks_ = 0;
    int* sub_blocks_ = nullptr;         // blocks completed per subsequence
    int* changed_ = nullptr;            // per-round change counters + grid barrier + status words
    int16_t *coef_ = nullptr, *dc_abs_ = nullptr;
    int* mcu_dc_ = nullptr;
    uint8_t *planes_ = nullptr, *frame_ = nullptr;
    void* tables_ = nullptr;            // JpegTables (device)
    int* pinned_status_ = nullptr;
    cudaEvent_t staged_ = nullptr;      // the pinned upload block is free again
    int last_launches_ = 0;
    size_t last_upload_ = 0;
    int max_coresident_ = 0;
    cudaEvent_t stage_ev_[kStages + 1] = {};
    bool profiling_ = false;
};

}  // namespace rmr
