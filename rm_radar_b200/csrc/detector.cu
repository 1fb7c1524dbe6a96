#include "detector.h"

#include <atomic>
#include <chrono>
#include <cstdio>

#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

namespace rmr {

namespace {

constexpr int kHeadOut = 64;     // detections per image copied with the counts; longer lists take a second copy
constexpr int kMaxOut = 1024;   // detections returned per image (the reference returns every survivor); more is a CapacityError

// NHWC4 fp16 -> planar float (test inspection: the blobKernel layout, detector.cu:151-171)
__global__ void input_to_planar_kernel(const __half* in, float* out, int hw, int n) {
    const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long>(n) * hw) return;
    const long img = i / hw, p = i % hw;
    for (int c = 0; c < 3; ++c) out[(img * 3 + c) * hw + p] = __half2float(in[i * 4 + c]);
}

struct DenseLevels {
    const float* logits[4];
    int h[4], w[4], stride[4], pitch[4], anchor0[4];
    int n_levels, anchors;
};

// Dense restatement of the exported Detect tail for *every* anchor, in the TensorRT output layout
// [n][4+nc][A] (cx, cy, w, h, class scores).  Test inspection only — the product path never
// materialises this tensor (decode_compact_kernel thresholds first).
__global__ void dense_tail_kernel(DenseLevels L, int nc, float* out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (a >= L.anchors) return;
    int l = 0;
    for (int i = 1; i < L.n_levels; ++i)
        if (a >= L.anchor0[i]) l = i;
    const int local = a - L.anchor0[l];
    const int px = local % L.w[l], py = local / L.w[l];
    const float* row = L.logits[l] + (static_cast<size_t>(img) * L.h[l] * L.w[l] + local) * L.pitch[l];
    float dist[4];
    for (int side = 0; side < 4; ++side) {
        float m = row[side * 16];
        for (int j = 1; j < 16; ++j) m = fmaxf(m, row[side * 16 + j]);
        float sum = 0.f, acc = 0.f;
        for (int j = 0; j < 16; ++j) {
            const float e = expf(row[side * 16 + j] - m);
            sum += e;
            acc += e * static_cast<float>(j);
        }
        dist[side] = acc / sum;
    }
    const float ax = px + 0.5f, ay = py + 0.5f, st = static_cast<float>(L.stride[l]);
    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    float* o = out + static_cast<size_t>(img) * (4 + nc) * L.anchors;
    o[0 * L.anchors + a] = (x1 + x2) * 0.5f * st;
    o[1 * L.anchors + a] = (y1 + y2) * 0.5f * st;
    o[2 * L.anchors + a] = (x2 - x1) * st;
    o[3 * L.anchors + a] = (y2 - y1) * st;
    for (int c = 0; c < nc; ++c) o[(4 + c) * L.anchors + a] = 1.f / (1.f + expf(-row[64 + c]));
}

}  // namespace

Detector::Detector(const std::string& engine_path, int classes, int image_w, int image_h, int max_batch,
                   float nms_thresh, float conf_thresh, int input_w, int input_h, bool compat, int device)
    : classes_(classes), image_w_(image_w), image_h_(image_h), max_batch_(max_batch), input_w_(input_w),
      input_h_(input_h), device_(device), nms_thresh_(nms_thresh), conf_thresh_(conf_thresh), compat_(compat) {
    if (max_batch <= 0) throw std::invalid_argument("max_batch_size must be positive");
    // `<x>.engine` / `<x>.onnx` / `<x>.rmeng`: load the plan, building it from the sibling ONNX file when it is not
    // there yet (detector.cpp:74-99); neither file -> std::invalid_argument before any device work, as there
    const std::string plan_path = resolve_engine(engine_path, input_h, input_w);
    RMR_CUDA(cudaSetDevice(device_));   // reference: cudaSetDevice(0) hard-coded (detector.cpp:61)
    cudaDeviceProp prop{};
    RMR_CUDA(cudaGetDeviceProperties(&prop, device_));
    if (prop.major != 10)
        throw CudaError(std::string("rm_radar_b200 needs an sm_100a device (B200); found ") + prop.name);
    net_ = std::make_unique<Net>(plan_path, max_batch);
    if (net_->in_w() != input_w || net_->in_h() != input_h)
        throw std::invalid_argument("engine input size does not match input_width/input_height");
    if (net_->num_classes() != classes)
        throw std::invalid_argument("engine class count (" + std::to_string(net_->num_classes()) +
                                    ") does not match `classes`");
    RMR_CUDA(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    post_alloc(post_, max_batch, kMaxOut);
    const size_t in_px = static_cast<size_t>(input_w) * input_h;
    RMR_CUDA(cudaMalloc(&staging_, in_px * 3 * max_batch));
    RMR_CUDA(cudaMemset(staging_, 0, in_px * 3 * max_batch));
    RMR_CUDA(cudaMalloc(&dev_geoms_, sizeof(LetterboxGeom) * max_batch));
    RMR_CUDA(cudaMallocHost(&pinned_geoms_, sizeof(LetterboxGeom) * max_batch));
    RMR_CUDA(cudaMallocHost(&pinned_out_, sizeof(Detection) * kMaxOut * max_batch));
    RMR_CUDA(cudaMallocHost(&pinned_counts_, sizeof(int) * max_batch));
    RMR_CUDA(cudaMallocHost(&pinned_cand_counts_, sizeof(int) * max_batch));
    // the NMS kernel writes its results into the pinned buffers itself (device-visible under unified addressing)
    RMR_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&post_.host_out), pinned_out_, 0));
    RMR_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&post_.host_out_count), pinned_counts_, 0));
    RMR_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&post_.host_cand_count), pinned_cand_counts_, 0));
    post_.host_head = kHeadOut;
    {
        int* done = nullptr;
        void* done_dev = nullptr;
        RMR_CUDA(cudaMallocHost(&done, sizeof(int) * max_batch));
        std::memset(done, 0, sizeof(int) * max_batch);
        RMR_CUDA(cudaHostGetDevicePointer(&done_dev, done, 0));
        pinned_done_ = done;
        post_.host_done = static_cast<int*>(done_dev);
    }
    RMR_CUDA(cudaEventCreate(&ev_fwd0_));
    RMR_CUDA(cudaEventCreate(&ev_fwd1_));
    frame_buffer(static_cast<size_t>(image_w) * image_h * 3);
    conv_init();
}

Detector::~Detector() {
    cudaSetDevice(device_);
    if (own_stream_) cudaStreamSynchronize(own_stream_);
    post_free(post_);
    if (ev_fwd0_) cudaEventDestroy(ev_fwd0_);
    if (ev_fwd1_) cudaEventDestroy(ev_fwd1_);
    cudaFree(staging_); cudaFree(dev_geoms_); cudaFree(dev_frame_);
    cudaFreeHost(pinned_geoms_); cudaFreeHost(pinned_out_); cudaFreeHost(pinned_counts_); cudaFreeHost(pinned_cand_counts_); cudaFreeHost(const_cast<int*>(pinned_done_)); cudaFreeHost(pinned_frame_);
    net_.reset();
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

uint8_t* Detector::frame_buffer(size_t bytes) {
    if (bytes > dev_frame_bytes_) {
        if (dev_frame_) RMR_CUDA(cudaFree(dev_frame_));
        if (pinned_frame_) RMR_CUDA(cudaFreeHost(pinned_frame_));
        RMR_CUDA(cudaMalloc(&dev_frame_, bytes));
        RMR_CUDA(cudaMallocHost(&pinned_frame_, bytes));
        dev_frame_bytes_ = pinned_frame_bytes_ = bytes;
    }
    return dev_frame_;
}

// enqueue(): everything up to the device->host copy of the survivors, asynchronously on stream_;
// collect(): the one synchronisation + unpacking.  run() = enqueue + collect; the cascade uses the split
// to overlap its own CPU work (and the Locator's launches) with the car network.
namespace {
const bool kTraceStages = [] { const char* e = std::getenv("RMR_TRACE"); return e && e[0] == '2'; }();
}

void Detector::enqueue(const uint8_t* dev_frame, int stride, const Roi* rois, int n) {
    pending_ = 0;
    const auto h0 = std::chrono::steady_clock::now();
    auto host_us = [&] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - h0).count(); };
    if (kTraceStages && !ev_trace_[0])
        for (auto& e : ev_trace_) RMR_CUDA(cudaEventCreate(&e));
    if (n == 0) return;   // Appendix B#8: the reference aborts inside TensorRT on an empty batch
    if (n > max_batch_) throw std::invalid_argument("batch larger than max_batch_size");
    RMR_CUDA(cudaSetDevice(device_));
    bool any_clean = false, any_unclean = false;
    for (int i = 0; i < n; ++i) {
        pinned_geoms_[i] = make_letterbox_geom(rois[i].x, rois[i].y, rois[i].w, rois[i].h, input_w_, input_h_, compat_);
        if (!pinned_geoms_[i].clean) ever_unclean_ = true;
    }
    // Bug-compatible mode keeps the persistent u8 staging buffer exactly like the reference once a
    // detector has produced (or can produce: ROI batches) non-640 geometry; a detector that only
    // ever sees clean geometry never exposes its staging bytes and takes the single fused kernel.
    const bool force_stage = compat_ && (ever_unclean_ || max_batch_ > 1);
    for (int i = 0; i < n; ++i) {
        if (force_stage) pinned_geoms_[i].clean = 0;
        (pinned_geoms_[i].clean ? any_clean : any_unclean) = true;
    }
    // same geometry as the last call (the car stage: one full frame of a fixed size): the device copy is still right
    if (uploaded_geoms_.size() != static_cast<size_t>(n) ||
        std::memcmp(uploaded_geoms_.data(), pinned_geoms_, sizeof(LetterboxGeom) * n) != 0) {
        RMR_CUDA(cudaMemcpyAsync(dev_geoms_, pinned_geoms_, sizeof(LetterboxGeom) * n, cudaMemcpyHostToDevice, stream_));
        uploaded_geoms_.assign(pinned_geoms_, pinned_geoms_ + n);
    }
    if (kTraceStages) RMR_CUDA(cudaEventRecord(ev_trace_[0], stream_));
    const double t_geom = host_us();
    launch_letterbox(dev_frame, stride, dev_geoms_, any_unclean, any_clean, n, staging_, net_->input(), input_w_,
                     input_h_, stream_);
    RMR_CUDA(cudaEventRecord(ev_fwd0_, stream_));
    const double t_lb = host_us();
    net_->forward(n, stream_);
    const double t_net = host_us();
    RMR_CUDA(cudaEventRecord(ev_fwd1_, stream_));
    post_.seq = ++seq_;   // what the NMS blocks of this call store into their completion flags
    launch_postprocess(net_->levels(), classes_, n, dev_geoms_, conf_thresh_, nms_thresh_, post_, stream_);
    if (kTraceStages) RMR_CUDA(cudaEventRecord(ev_trace_[1], stream_));
    // counters and the first kHeadOut survivors of every image are already on their way: nms_restore_kernel writes them
    // into the pinned buffers (PostBuffers::host_*); a longer list is copied in collect()
    if (kTraceStages) {
        RMR_CUDA(cudaEventRecord(ev_trace_[2], stream_));
        std::fprintf(stderr, "  enqueue(n=%d) host us: geoms %.1f | letterbox launched %.1f | graph launched %.1f | all enqueued %.1f\n", n,
                     t_geom, t_lb, t_net, host_us());
    }
    int net_launches = 0;
    net_->plan_stats(n, &net_launches, nullptr, nullptr);
    last_launches_ = (any_clean ? 1 : 0) + (any_unclean ? letterbox_compat_launches() : 0) + net_launches + 2;
    pending_ = n;
}

std::vector<std::vector<Detection>> Detector::collect() {
    const int n = pending_;
    pending_ = 0;
    std::vector<std::vector<Detection>> results(n);
    if (n == 0) return results;
    RMR_CUDA(cudaSetDevice(device_));
    // The results are written into pinned memory by the NMS kernel itself, one completion flag per image last: watching
    // the flags returns a few microseconds after the last write, a stream synchronise only after the kernel has retired
    // and the driver has noticed.  Bounded: after ~2 ms of watching (or with RMR_SYNC_WAIT=1) fall back to the synchronise,
    // which also surfaces any launch failure.
    static const bool sync_wait = [] { const char* e = std::getenv("RMR_SYNC_WAIT"); return e && e[0] == '1'; }();
    bool seen = false;
    if (!sync_wait) {
        const auto t0 = std::chrono::steady_clock::now();
        for (long spins = 0; !seen; ++spins) {
            seen = true;
            for (int i = 0; i < n; ++i) seen = seen && (pinned_done_[i] == seq_);
            if (!seen && (spins & 1023) == 1023 &&
                std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (!seen) RMR_CUDA(cudaStreamSynchronize(stream_));
    if (cudaEventElapsedTime(&last_forward_ms_, ev_fwd0_, ev_fwd1_) != cudaSuccess) {
        (void)cudaGetLastError();   // cudaErrorNotReady is not an error of ours
        RMR_CUDA(cudaStreamSynchronize(stream_));
        RMR_CUDA(cudaEventElapsedTime(&last_forward_ms_, ev_fwd0_, ev_fwd1_));
    }
    if (kTraceStages) {
        float lb = 0, post = 0, d2h = 0;
        cudaEventElapsedTime(&lb, ev_trace_[0], ev_fwd0_);
        cudaEventElapsedTime(&post, ev_fwd1_, ev_trace_[1]);
        cudaEventElapsedTime(&d2h, ev_trace_[1], ev_trace_[2]);
        std::fprintf(stderr, "  collect(n=%d) device us: letterbox %.1f | net %.1f | decode+nms %.1f | d2h %.1f\n", n, lb * 1e3,
                     last_forward_ms_ * 1e3, post * 1e3, d2h * 1e3);
    }
    int longest = 0;
    for (int i = 0; i < n; ++i) {
        // the reference returns every survivor; a truncated list would be a different result, so it is an error
        if (pinned_cand_counts_[i] > kMaxCandidates)
            throw CapacityError("image " + std::to_string(i) + ": " + std::to_string(pinned_cand_counts_[i]) +
                                " anchors pass the confidence threshold, capacity " + std::to_string(kMaxCandidates));
        if (pinned_counts_[i] > kMaxOut)
            throw CapacityError("image " + std::to_string(i) + ": " + std::to_string(pinned_counts_[i]) +
                                " detections survive NMS, capacity " + std::to_string(kMaxOut));
        longest = std::max(longest, pinned_counts_[i]);
    }
    if (longest > kHeadOut) {
        RMR_CUDA(cudaMemcpyAsync(pinned_out_, post_.out, sizeof(Detection) * kMaxOut * n, cudaMemcpyDeviceToHost, stream_));
        RMR_CUDA(cudaStreamSynchronize(stream_));
    }
    for (int i = 0; i < n; ++i) {
        const int c = std::min(pinned_counts_[i], kMaxOut);
        results[i].assign(pinned_out_ + static_cast<size_t>(i) * kMaxOut, pinned_out_ + static_cast<size_t>(i) * kMaxOut + c);
    }
    return results;
}

std::vector<std::vector<Detection>> Detector::run(const uint8_t* dev_frame, int stride, const Roi* rois, int n) {
    enqueue(dev_frame, stride, rois, n);
    return collect();
}

std::vector<Detection> Detector::detect_host(const uint8_t* bgr, int w, int h, int stride) {
    if (!bgr || w <= 0 || h <= 0 || stride < w * 3) throw std::invalid_argument("bad image");
    RMR_CUDA(cudaSetDevice(device_));
    const size_t bytes = static_cast<size_t>(w) * h * 3;
    uint8_t* dev = frame_buffer(bytes);
    // reference: memcpy into pinned+mapped host memory, kernels read it over PCIe (detector.cu:388-400);
    // here: one pinned staging copy + one async DMA, the frame then stays resident for the ROI stage
    for (int y = 0; y < h; ++y) std::memcpy(pinned_frame_ + static_cast<size_t>(y) * w * 3, bgr + static_cast<size_t>(y) * stride, static_cast<size_t>(w) * 3);
    RMR_CUDA(cudaMemcpyAsync(dev, pinned_frame_, bytes, cudaMemcpyHostToDevice, stream_));
    const Roi roi{0, 0, w, h};
    return run(dev, w * 3, &roi, 1)[0];
}

std::vector<std::vector<Detection>> Detector::detect_host_batch(const uint8_t* const* bgr, const int* w, const int* h,
                                                                const int* stride, int n) {
    std::vector<std::vector<Detection>> results;
    if (n <= 0) return results;
    if (n > max_batch_) throw std::invalid_argument("batch larger than max_batch_size");
    RMR_CUDA(cudaSetDevice(device_));
    // pack the images as one tall strip of the widest row pitch so that a single frame pointer +
    // stride serves every ROI (the reference packs them back to back in image_ptr_, detector.cu:456)
    int maxw = 0;
    long rows = 0;
    for (int i = 0; i < n; ++i) {
        if (!bgr[i] || w[i] <= 0 || h[i] <= 0) throw std::invalid_argument("bad image in batch");
        maxw = std::max(maxw, w[i]);
        rows += h[i];
    }
    const size_t pitch = static_cast<size_t>(maxw) * 3;
    uint8_t* dev = frame_buffer(pitch * rows);
    std::vector<Roi> rois(n);
    long row0 = 0;
    for (int i = 0; i < n; ++i) {
        for (int y = 0; y < h[i]; ++y)
            std::memcpy(pinned_frame_ + (row0 + y) * pitch, bgr[i] + static_cast<size_t>(y) * stride[i], static_cast<size_t>(w[i]) * 3);
        rois[i] = Roi{0, static_cast<int>(row0), w[i], h[i]};
        row0 += h[i];
    }
    RMR_CUDA(cudaMemcpyAsync(dev, pinned_frame_, pitch * rows, cudaMemcpyHostToDevice, stream_));
    return run(dev, static_cast<int>(pitch), rois.data(), n);
}

std::vector<std::vector<Detection>> Detector::detect_device_rois(const uint8_t* dev_frame, int stride,
                                                                 const Roi* rois, int n) {
    return run(dev_frame, stride, rois, n);
}

void Detector::last_input(float* out, int n) {
    RMR_CUDA(cudaSetDevice(device_));
    const int hw = input_w_ * input_h_;
    float* dev = nullptr;
    RMR_CUDA(cudaMalloc(&dev, sizeof(float) * 3 * hw * n));
    const long total = static_cast<long>(n) * hw;
    input_to_planar_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, stream_>>>(net_->input(), dev, hw, n);
    RMR_CUDA(cudaMemcpyAsync(out, dev, sizeof(float) * 3 * hw * n, cudaMemcpyDeviceToHost, stream_));
    RMR_CUDA(cudaStreamSynchronize(stream_));
    cudaFree(dev);
}

void Detector::last_output(float* out, int n) {
    RMR_CUDA(cudaSetDevice(device_));
    DenseLevels L{};
    const auto& lv = net_->levels();
    L.n_levels = static_cast<int>(lv.size());
    int a0 = 0;
    for (int i = 0; i < L.n_levels; ++i) {
        L.logits[i] = lv[i].logits; L.h[i] = lv[i].h; L.w[i] = lv[i].w; L.stride[i] = lv[i].stride;
        L.pitch[i] = lv[i].pitch; L.anchor0[i] = a0;
        a0 += lv[i].h * lv[i].w;
    }
    L.anchors = a0;
    const size_t count = static_cast<size_t>(n) * (4 + classes_) * a0;
    float* dev = nullptr;
    RMR_CUDA(cudaMalloc(&dev, sizeof(float) * count));
    dense_tail_kernel<<<dim3((a0 + 127) / 128, n), 128, 0, stream_>>>(L, classes_, dev);
    RMR_CUDA(cudaMemcpyAsync(out, dev, sizeof(float) * count, cudaMemcpyDeviceToHost, stream_));
    RMR_CUDA(cudaStreamSynchronize(stream_));
    cudaFree(dev);
}

// ------------------------------------------------------------------------------------------
// RobotDetector
// ------------------------------------------------------------------------------------------
namespace {

// computeIoU — detector.cpp:324-349: intersection / area of the *bounding* rectangle (Appendix B#16)
float compute_iou_bounding(const float* a, const float* b) {
    float x1 = std::max(a[0], b[0]), y1 = std::max(a[1], b[1]);
    float x2 = std::min(a[0] + a[2], b[0] + b[2]), y2 = std::min(a[1] + a[3], b[1] + b[3]);
    float iw = 0.f, ih = 0.f;
    if (x1 < x2 && y1 < y2) { iw = x2 - x1; ih = y2 - y1; }
    x1 = std::min(a[0], b[0]); y1 = std::min(a[1], b[1]);
    x2 = std::max(a[0] + a[2], b[0] + b[2]); y2 = std::max(a[1] + a[3], b[1] + b[3]);
    const float ia = iw * ih, ua = (x2 - x1) * (y2 - y1);
    return ua > 0 ? ia / ua : 0.f;
}

// Robot::rect(): optional<Rect2f> -> optional<cv::Rect> = cvRound per field (robot.h:111)
void rect_rounded(const float* r, float* out) {
    for (int i = 0; i < 4; ++i) out[i] = static_cast<float>(std::lrintf(r[i]));
}

// Robot::setDetection — robot.cpp:41-74
RobotRecord set_detection(const Detection& car, const std::vector<Detection>& armors) {
    RobotRecord r;
    r.rect[0] = car.x; r.rect[1] = car.y; r.rect[2] = car.width; r.rect[3] = car.height;
    r.has_rect = true;
    if (armors.empty()) return r;
    std::map<int, float> score;
    for (const Detection& a : armors) score[static_cast<int>(a.label)] += a.confidence;
    int label = score.begin()->first;
    float best = score.begin()->second;
    for (const auto& kv : score)
        if (best < kv.second) { best = kv.second; label = kv.first; }
    int count = 0;
    for (const Detection& a : armors) count += (a.label == static_cast<float>(label)) ? 1 : 0;
    r.label = label;
    r.confidence = best / static_cast<float>(count);
    r.armors = armors;
    for (Detection& a : r.armors) { a.x += car.x; a.y += car.y; }
    r.detected = true;
    return r;
}

}  // namespace

RobotDetector::RobotDetector(const std::string& car_engine, const std::string& armor_engine, int image_w,
                             int image_h, int armor_classes, int max_cars, float iou_thresh, float car_nms,
                             float car_conf, float armor_nms, float armor_conf, int input_w, int input_h, bool compat,
                             int device, int frames)
    : max_cars_(max_cars), iou_thresh_(iou_thresh), frames_(frames) {
    if (frames < 1 || frames > 64) throw std::invalid_argument("frames must be in [1, 64]");
    // detector.cpp:377-404: car detector batch 1 / 1 class, armor detector batch max_cars.  Throughput mode: `frames`
    // images per car-network call; the armor network takes the ROIs of all of them in chunks of its batch size
    car_ = std::make_unique<Detector>(car_engine, 1, image_w, image_h * frames, frames, car_nms, car_conf, input_w, input_h, compat, device);
    const int armor_batch = frames == 1 ? max_cars : std::min(max_cars * frames, std::max(max_cars, 128));
    armor_ = std::make_unique<Detector>(armor_engine, armor_classes, image_w, image_h * frames, armor_batch, armor_nms, armor_conf,
                                        input_w, input_h, compat, device);
    armor_->set_stream(car_->stream());
    RMR_CUDA(cudaEventCreateWithFlags(&ev_frame_, cudaEventDisableTiming));
}

RobotDetector::~RobotDetector() {
    if (ev_frame_) cudaEventDestroy(ev_frame_);
}

// begin(): upload (host frames) and enqueue the car stage, nothing waits; finish(): the rest of
// RobotDetector::detect (detector.cpp:413-455).  detect_host / detect_device = begin + finish.
void RobotDetector::begin(const uint8_t* frame, bool on_device, int w, int h, int stride) {
    if (!frame || w <= 0 || h <= 0 || stride < w * 3) throw std::invalid_argument("bad image");
    RMR_CUDA(cudaSetDevice(car_->device()));
    const uint8_t* dev = frame;
    int dev_stride = stride;
    if (!on_device) {
        const size_t bytes = static_cast<size_t>(w) * h * 3;
        uint8_t* buf = car_->frame_buffer(bytes);
        if (stride == w * 3) {
            RMR_CUDA(cudaMemcpyAsync(buf, frame, bytes, cudaMemcpyHostToDevice, car_->stream()));
        } else {
            RMR_CUDA(cudaMemcpy2DAsync(buf, static_cast<size_t>(w) * 3, frame, stride, static_cast<size_t>(w) * 3, h,
                                       cudaMemcpyHostToDevice, car_->stream()));
        }
        dev = buf;
        dev_stride = w * 3;
        RMR_CUDA(cudaEventRecord(ev_frame_, car_->stream()));
    }
    cur_frame_ = dev; cur_w_ = w; cur_h_ = h; cur_stride_ = dev_stride;
    mid_done_ = false;
    const Roi full{0, 0, w, h};
    car_->enqueue(dev, dev_stride, &full, 1);
}

std::vector<RobotRecord> RobotDetector::detect_host(const uint8_t* bgr, int w, int h, int stride) {
    begin(bgr, false, w, h, stride);
    return finish();
}

std::vector<RobotRecord> RobotDetector::detect_device(const uint8_t* dev_bgr, int w, int h, int stride) {
    begin(dev_bgr, true, w, h, stride);
    return finish();
}

const std::vector<Detection>& RobotDetector::cars() {
    const uint8_t* dev_bgr = cur_frame_;
    const int stride = cur_stride_;
    if (dev_bgr == nullptr) throw std::invalid_argument("RobotDetector::finish without begin");
    if (mid_done_) return last_cars_;
    std::vector<Detection> cars;
    try {
        cars = std::move(car_->collect()[0]);
    } catch (...) {
        cur_frame_ = nullptr;      // the frame is over: the next call starts with begin()
        throw;
    }
    last_launches_ = car_->last_launches();
    last_flops_ = car_->net().flops_per_image();
    last_car_ms_ = car_->last_forward_ms();
    last_armor_ms_ = 0.f;
    // Appendix B#8: more cars than max_batch_size is UB in the reference; keep the first max_cars
    if (static_cast<int>(cars.size()) > max_cars_) cars.resize(max_cars_);
    // cv::Rect(float, float, float, float): truncation (detector.cpp:420-421); ROIs are read from the
    // resident frame instead of `image(rect).clone()`
    std::vector<Roi> rois;
    roi_of_car_.assign(cars.size(), -1);
    for (size_t i = 0; i < cars.size(); ++i) {
        Roi r{static_cast<int>(cars[i].x), static_cast<int>(cars[i].y), static_cast<int>(cars[i].width),
              static_cast<int>(cars[i].height)};
        if (r.w <= 0 || r.h <= 0) continue;   // cv::Mat ROI of zero area: nothing to detect in
        roi_of_car_[i] = static_cast<int>(rois.size());
        rois.push_back(r);
    }
    if (!rois.empty()) {
        armor_->enqueue(dev_bgr, stride, rois.data(), static_cast<int>(rois.size()));
        last_launches_ += armor_->last_launches();
        last_flops_ += armor_->net().flops_per_image() * rois.size();
    }
    last_cars_ = std::move(cars);
    mid_done_ = true;
    return last_cars_;
}

std::vector<RobotRecord> RobotDetector::finish() {
    cars();
    cur_frame_ = nullptr;
    mid_done_ = false;
    std::vector<std::vector<Detection>> armor_batch = armor_->collect();   // empty when nothing was enqueued
    if (!armor_batch.empty()) last_armor_ms_ = armor_->last_forward_ms();
    last_armors_.assign(last_cars_.size(), {});
    for (size_t i = 0; i < last_cars_.size(); ++i)
        if (roi_of_car_[i] >= 0) last_armors_[i] = armor_batch[roi_of_car_[i]];
    return assemble(last_cars_, last_armors_);
}

// Robot::setDetection per car, then the label de-duplication of RobotDetector::detect (detector.cpp:426-455)
std::vector<RobotRecord> RobotDetector::assemble(const std::vector<Detection>& cars,
                                                 const std::vector<std::vector<Detection>>& armors) {
    std::vector<RobotRecord> robots;
    robots.reserve(cars.size());
    std::map<int, RobotRecord> by_label;
    for (size_t i = 0; i < cars.size(); ++i) {
        RobotRecord robot = set_detection(cars[i], armors[i]);
        robot.car = static_cast<int>(i);
        if (!robot.detected) {
            robots.push_back(robot);
            continue;
        }
        auto it = by_label.find(robot.label);
        if (it == by_label.end()) {
            by_label.emplace(robot.label, robot);
        } else {
            float ra[4], rb[4];
            rect_rounded(it->second.rect, ra);
            rect_rounded(robot.rect, rb);
            if (compute_iou_bounding(ra, rb) > iou_thresh_) continue;
            if (it->second.confidence < robot.confidence) it->second = robot;
        }
    }
    for (auto& kv : by_label) robots.push_back(kv.second);
    return robots;
}

// ---- throughput mode: `n` frames per call ------------------------------------------------------------------
void RobotDetector::begin_batch(const uint8_t* frames, bool on_device, int n, int w, int h, int stride) {
    if (!frames || n <= 0 || w <= 0 || h <= 0 || stride < w * 3) throw std::invalid_argument("bad image batch");
    if (n > frames_) throw std::invalid_argument("more frames than the detector was created for");
    RMR_CUDA(cudaSetDevice(car_->device()));
    const uint8_t* dev = frames;
    if (!on_device) {
        const size_t bytes = static_cast<size_t>(stride) * h * n;
        uint8_t* buf = car_->frame_buffer(bytes);
        RMR_CUDA(cudaMemcpyAsync(buf, frames, bytes, cudaMemcpyHostToDevice, car_->stream()));
        RMR_CUDA(cudaEventRecord(ev_frame_, car_->stream()));
        dev = buf;
    }
    cur_frame_ = dev; cur_w_ = w; cur_h_ = h; cur_stride_ = stride; cur_n_ = n;
    mid_done_ = false;
    std::vector<Roi> full(n);
    for (int i = 0; i < n; ++i) full[i] = Roi{0, i * h, w, h};   // the batch is one tall strip of rows
    car_->enqueue(dev, stride, full.data(), n);
}

const std::vector<std::vector<Detection>>& RobotDetector::batch_cars() {
    if (cur_frame_ == nullptr) throw std::invalid_argument("RobotDetector::finish_batch without begin_batch");
    if (mid_done_) return batch_cars_;
    const int n = cur_n_, h = cur_h_;
    try {
        batch_cars_ = car_->collect();
    } catch (...) {
        cur_frame_ = nullptr;
        throw;
    }
    last_launches_ = car_->last_launches();
    last_flops_ = car_->net().flops_per_image() * n;
    last_car_ms_ = car_->last_forward_ms();
    last_armor_ms_ = 0.f;
    // ROIs of every frame, in frame order; (frame, car) of each ROI
    batch_rois_.clear();
    batch_roi_of_car_.assign(n, {});
    for (int f = 0; f < n; ++f) {
        std::vector<Detection>& cars = batch_cars_[f];
        if (static_cast<int>(cars.size()) > max_cars_) cars.resize(max_cars_);
        batch_roi_of_car_[f].assign(cars.size(), -1);
        for (size_t i = 0; i < cars.size(); ++i) {
            const Detection& c = cars[i];
            Roi r{static_cast<int>(c.x), static_cast<int>(c.y) + f * h, static_cast<int>(c.width), static_cast<int>(c.height)};
            if (r.w <= 0 || r.h <= 0) continue;
            batch_roi_of_car_[f][i] = static_cast<int>(batch_rois_.size());
            batch_rois_.push_back(r);
        }
    }
    // the first armor chunk goes out now, so the host work that follows (the searches) overlaps it
    if (!batch_rois_.empty()) {
        const int m = static_cast<int>(std::min<size_t>(armor_->max_batch(), batch_rois_.size()));
        armor_->enqueue(cur_frame_, cur_stride_, batch_rois_.data(), m);
    }
    mid_done_ = true;
    return batch_cars_;
}

std::vector<std::vector<RobotRecord>> RobotDetector::finish_batch() {
    batch_cars();
    const uint8_t* dev = cur_frame_;
    const int n = cur_n_, stride = cur_stride_;
    cur_frame_ = nullptr;
    mid_done_ = false;
    const std::vector<std::vector<Detection>>& cars = batch_cars_;
    const std::vector<Roi>& rois = batch_rois_;
    const std::vector<std::vector<int>>& roi_of_car = batch_roi_of_car_;
    std::vector<std::vector<Detection>> armor_all(rois.size());
    const int chunk = armor_->max_batch();
    for (size_t r0 = 0; r0 < rois.size(); r0 += chunk) {
        const int m = static_cast<int>(std::min<size_t>(chunk, rois.size() - r0));
        auto part = r0 == 0 ? armor_->collect() : armor_->detect_device_rois(dev, stride, rois.data() + r0, m);
        for (int i = 0; i < m; ++i) armor_all[r0 + i] = std::move(part[i]);
        last_launches_ += armor_->last_launches();
        last_armor_ms_ += armor_->last_forward_ms();
        last_flops_ += armor_->net().flops_per_image() * m;
    }
    std::vector<std::vector<RobotRecord>> out(n);
    for (int f = 0; f < n; ++f) {
        std::vector<std::vector<Detection>> armors(cars[f].size());
        for (size_t i = 0; i < cars[f].size(); ++i)
            if (roi_of_car[f][i] >= 0) armors[i] = armor_all[roi_of_car[f][i]];
        out[f] = assemble(cars[f], armors);
        if (f == n - 1) { last_cars_ = cars[f]; last_armors_ = armors; }
    }
    return out;
}

}  // namespace rmr
