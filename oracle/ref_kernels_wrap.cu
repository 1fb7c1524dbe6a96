// ORACLE (test infrastructure only).  Host-side launchers around the reference's OWN CUDA kernels
// (resizeKernel, copyMakeBorderKernel, blobKernel, transposeKernel, decodeKernel, IoU, NMSKernel —
// /root/reference/src/detect/detector.cu:40-360).  Those kernels depend on nothing but
// cuda_runtime.h and `Detection`, so the recipe in oracle/Makefile lifts that namespace block verbatim
// into the git-ignored oracle/_ref/ref_kernels.inc (never into the repository) and this file compiles
// it with nvcc for sm_100a into oracle/_ref/libref_kernels.so.  tests/test_gpu_ref_kernels.py runs the
// reference kernels on the B200 with the launch shapes of the reference's call sites
// (detector.cu:380-421, 522-548) and pins oracle/detect_oracle.py to their outputs.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#include "detection.h"        // oracle/_ref/detection.h (copied from the reference by the recipe)
using radar::Detection;
#include "ref_kernels.inc"    // oracle/_ref/ref_kernels.inc = namespace radar::detect { ...kernels... }

using namespace radar::detect;

namespace {
template <typename T>
struct DevBuf {
    T* p = nullptr;
    explicit DevBuf(size_t n) { cudaMalloc(&p, n * sizeof(T)); }
    ~DevBuf() { cudaFree(p); }
};
int done() {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    return e == cudaSuccess ? 0 : static_cast<int>(e);
}
}  // namespace

extern "C" {

// detector.cu:392-400: block 16x16, grid over the destination
int ref_resize(const unsigned char* src, unsigned char* dst, int channels, int src_w, int src_h, int dst_w, int dst_h) {
    DevBuf<unsigned char> s(static_cast<size_t>(src_w) * src_h * channels), d(static_cast<size_t>(dst_w) * dst_h * channels);
    cudaMemcpy(s.p, src, static_cast<size_t>(src_w) * src_h * channels, cudaMemcpyHostToDevice);
    dim3 block(16, 16), grid((dst_w + 15) / 16, (dst_h + 15) / 16);
    resizeKernel<<<grid, block>>>(s.p, d.p, channels, src_w, src_h, dst_w, dst_h);
    cudaMemcpy(dst, d.p, static_cast<size_t>(dst_w) * dst_h * channels, cudaMemcpyDeviceToHost);
    return done();
}

// detector.cu:402-410: the destination buffer is persistent in the reference (stale bytes survive where the
// kernel does not write), so the caller passes its previous contents in `dst`; grid covers grid_w x grid_h
int ref_copy_make_border(const unsigned char* src, unsigned char* dst, int channels, int src_w, int src_h, int top, int bottom,
                         int left, int right, int grid_w, int grid_h, int dst_bytes) {
    DevBuf<unsigned char> s(static_cast<size_t>(src_w) * src_h * channels), d(static_cast<size_t>(dst_bytes));
    cudaMemcpy(s.p, src, static_cast<size_t>(src_w) * src_h * channels, cudaMemcpyHostToDevice);
    cudaMemcpy(d.p, dst, static_cast<size_t>(dst_bytes), cudaMemcpyHostToDevice);
    dim3 block(16, 16), grid((grid_w + 15) / 16, (grid_h + 15) / 16);
    copyMakeBorderKernel<<<grid, block>>>(s.p, d.p, channels, src_w, src_h, top, bottom, left, right);
    cudaMemcpy(dst, d.p, static_cast<size_t>(dst_bytes), cudaMemcpyDeviceToHost);
    return done();
}

// detector.cu:412-414
int ref_blob(const unsigned char* src, float* dst, int width, int height, int channels, float scale) {
    const size_t n = static_cast<size_t>(width) * height * channels;
    DevBuf<unsigned char> s(n);
    DevBuf<float> d(n);
    cudaMemcpy(s.p, src, n, cudaMemcpyHostToDevice);
    dim3 block(16, 16), grid((width + 15) / 16, (height + 15) / 16);
    blobKernel<<<grid, block>>>(s.p, d.p, width, height, channels, scale);
    cudaMemcpy(dst, d.p, n * sizeof(float), cudaMemcpyDeviceToHost);
    return done();
}

// detector.cu:528-534: src [rows][cols] -> dst [cols][rows]
int ref_transpose(const float* src, float* dst, int rows, int cols) {
    const size_t n = static_cast<size_t>(rows) * cols;
    DevBuf<float> s(n), d(n);
    cudaMemcpy(s.p, src, n * sizeof(float), cudaMemcpyHostToDevice);
    dim3 block(32, 32), grid((cols + 31) / 32, (rows + 31) / 32);
    transposeKernel<<<grid, block>>>(s.p, d.p, rows, cols);
    cudaMemcpy(dst, d.p, n * sizeof(float), cudaMemcpyDeviceToHost);
    return done();
}

// detector.cu:536-540: src [anchors][channels] (already transposed), dst [anchors][6]
int ref_decode(const float* src, float* dst, int channels, int anchors, int classes) {
    DevBuf<float> s(static_cast<size_t>(anchors) * channels), d(static_cast<size_t>(anchors) * 6);
    cudaMemcpy(s.p, src, static_cast<size_t>(anchors) * channels * sizeof(float), cudaMemcpyHostToDevice);
    decodeKernel<<<(anchors + 31) / 32, 32>>>(s.p, d.p, channels, anchors, classes);
    cudaMemcpy(dst, d.p, static_cast<size_t>(anchors) * 6 * sizeof(float), cudaMemcpyDeviceToHost);
    return done();
}

// detector.cu:542-548: in place on [anchors][6]; suppressed / below-threshold rows get label = NaN
int ref_nms(float* det, float nms_thresh, float score_thresh, int anchors) {
    DevBuf<float> d(static_cast<size_t>(anchors) * 6);
    cudaMemcpy(d.p, det, static_cast<size_t>(anchors) * 6 * sizeof(float), cudaMemcpyHostToDevice);
    const int bs = 16 * 16;
    dim3 grid((anchors + bs - 1) / bs, (anchors + bs - 1) / bs);
    NMSKernel<<<grid, bs, bs * sizeof(Detection)>>>(d.p, nms_thresh, score_thresh, anchors);
    cudaMemcpy(det, d.p, static_cast<size_t>(anchors) * 6 * sizeof(float), cudaMemcpyDeviceToHost);
    return done();
}

// the reference's IoU on the host (detector.cu:270-295 is __host__ __device__)
float ref_iou(float x1, float y1, float w1, float h1, float x2, float y2, float w2, float h2) {
    return IoU(x1, y1, w1, h1, x2, y2, w2, h2);
}

}  // extern "C"
