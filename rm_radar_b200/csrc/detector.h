// Detector (one network) and RobotDetector (car -> armor cascade) on top of Net + the fused
// pre/post-process kernels.  Mirrors radar::Detector / radar::RobotDetector
// (/root/reference/src/detect/detector.h:84-190, detector.cpp:377-455).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "net.h"
#include "postprocess.h"
#include "preprocess.h"

namespace rmr {

struct Roi {
    int x, y, w, h;
};

class Detector {
public:
    Detector(const std::string& engine_path, int classes, int image_w, int image_h, int max_batch, float nms_thresh,
             float conf_thresh, int input_w, int input_h, bool compat, int device);
    ~Detector();

    // single host image (Detector::detect(const cv::Mat&))
    std::vector<Detection> detect_host(const uint8_t* bgr, int w, int h, int stride);
    // batch of host images (Detector::detect(container))
    std::vector<std::vector<Detection>> detect_host_batch(const uint8_t* const* bgr, const int* w, const int* h,
                                                          const int* stride, int n);
    // ROIs of a frame that already lives in device memory (the cascade's second stage)
    std::vector<std::vector<Detection>> detect_device_rois(const uint8_t* dev_frame, int stride, const Roi* rois,
                                                           int n);
    // asynchronous halves of detect_device_rois (same stream): enqueue everything, then wait + unpack
    void enqueue(const uint8_t* dev_frame, int stride, const Roi* rois, int n);
    std::vector<std::vector<Detection>> collect();
    void last_input(float* out, int n);    // [n][3][H][W] float
    void last_output(float* out, int n);   // [n][4+nc][A] float
    void set_stream(cudaStream_t s) { stream_ = s; }
    cudaStream_t stream() const { return stream_; }
    Net& net() { return *net_; }
    int classes() const { return classes_; }
    int max_batch() const { return max_batch_; }
    int device() const { return device_; }
    uint8_t* frame_buffer(size_t bytes);   // device staging for host frames (grows on demand)
    int last_launches() const { return last_launches_; }
    // device time of the last call's network replay (CUDA events on the detector's stream around Net::forward)
    float last_forward_ms() const { return last_forward_ms_; }

private:
    std::vector<std::vector<Detection>> run(const uint8_t* dev_frame, int stride, const Roi* rois, int n);
    int pending_ = 0;

    int classes_, image_w_, image_h_, max_batch_, input_w_, input_h_, device_;
    float nms_thresh_, conf_thresh_;
    bool compat_, ever_unclean_ = false;
    cudaStream_t stream_ = nullptr, own_stream_ = nullptr;
    std::unique_ptr<Net> net_;
    PostBuffers post_;
    uint8_t* dev_frame_ = nullptr;
    size_t dev_frame_bytes_ = 0;
    uint8_t* pinned_frame_ = nullptr;
    size_t pinned_frame_bytes_ = 0;
    uint8_t* staging_ = nullptr;            // the reference's dev_border_ptr_: [max_batch][H*W*3] u8
    LetterboxGeom *dev_geoms_ = nullptr, *pinned_geoms_ = nullptr;
    std::vector<LetterboxGeom> uploaded_geoms_;   // what dev_geoms_ holds
    Detection* pinned_out_ = nullptr;
    int* pinned_counts_ = nullptr;
    int* pinned_cand_counts_ = nullptr;    // candidates per image before NMS (capacity check)
    volatile int* pinned_done_ = nullptr;  // completion flag per image, written by the NMS kernel (PostBuffers::host_done)
    int seq_ = 0;
    int last_launches_ = 0;
    cudaEvent_t ev_fwd0_ = nullptr, ev_fwd1_ = nullptr;
    cudaEvent_t ev_trace_[3] = {nullptr, nullptr, nullptr};   // RMR_TRACE=2: stage boundaries on the stream
    float last_forward_ms_ = 0.f;
};

struct RobotRecord {
    float rect[4];
    bool has_rect = false, detected = false;
    int label = -1;
    float confidence = 0.f;
    std::vector<Detection> armors;
    int car = -1;           // index of the car detection this robot was made from (rect == that car's box)
};

class RobotDetector {
public:
    // frames > 1: throughput mode — the car network runs `frames` images per call and the armor network every
    // ROI of those frames (Detector::detect(container), /root/reference/src/detect/detector.cu:439-502)
    RobotDetector(const std::string& car_engine, const std::string& armor_engine, int image_w, int image_h,
                  int armor_classes, int max_cars, float iou_thresh, float car_nms, float car_conf, float armor_nms,
                  float armor_conf, int input_w, int input_h, bool compat, int device, int frames = 1);
    ~RobotDetector();
    // frames of one size, back to back in memory (frame i at frame + i * h * stride); begin_batch enqueues the car
    // stage for all of them and returns at once, finish_batch does the rest and returns the robots per frame
    void begin_batch(const uint8_t* frames, bool on_device, int n, int w, int h, int stride);
    std::vector<std::vector<RobotRecord>> finish_batch();
    int frames() const { return frames_; }
    // recorded on the detector stream right after the host->device copy of the frame(s) of the current call: other
    // uploads (the LiDAR cloud) queue behind it, so the frame — the one copy on the critical path — has the link to itself
    cudaEvent_t frame_uploaded() const { return ev_frame_; }
    std::vector<RobotRecord> detect_host(const uint8_t* bgr, int w, int h, int stride);
    std::vector<RobotRecord> detect_device(const uint8_t* dev_bgr, int w, int h, int stride);
    // split form: begin() uploads / enqueues the car stage and returns at once; cars() waits for the car stage,
    // enqueues the armor stage and returns the car boxes (what Locator::search needs — it can run beside the armor
    // network); finish() waits for the armor stage and assembles the robots (calls cars() itself if nobody did)
    void begin(const uint8_t* frame, bool on_device, int w, int h, int stride);
    const std::vector<Detection>& cars();
    std::vector<RobotRecord> finish();
    // same split for a batch: boxes of every frame's cars, in frame order
    const std::vector<std::vector<Detection>>& batch_cars();
    const std::vector<std::vector<Detection>>& last_batch_cars() const { return batch_cars_; }
    void set_stream(cudaStream_t s) { car_->set_stream(s); armor_->set_stream(s); }
    Detector& car() { return *car_; }
    Detector& armor() { return *armor_; }
    const std::vector<Detection>& last_cars() const { return last_cars_; }
    const std::vector<std::vector<Detection>>& last_armors() const { return last_armors_; }
    int last_launches() const { return last_launches_; }
    double last_flops() const { return last_flops_; }
    float last_car_ms() const { return last_car_ms_; }
    float last_armor_ms() const { return last_armor_ms_; }

private:
    std::unique_ptr<Detector> car_, armor_;
    int max_cars_;
    float iou_thresh_;
    std::vector<Detection> last_cars_;
    std::vector<std::vector<Detection>> last_armors_;
    int last_launches_ = 0;
    double last_flops_ = 0;
    float last_car_ms_ = 0.f, last_armor_ms_ = 0.f;
    const uint8_t* cur_frame_ = nullptr;
    int cur_w_ = 0, cur_h_ = 0, cur_stride_ = 0, cur_n_ = 0;
    int frames_ = 1;
    cudaEvent_t ev_frame_ = nullptr;
    bool mid_done_ = false;
    std::vector<int> roi_of_car_;                       // single frame: armor batch slot of car i, -1 = no ROI
    std::vector<std::vector<Detection>> batch_cars_;    // batch: cars per frame
    std::vector<Roi> batch_rois_;
    std::vector<std::vector<int>> batch_roi_of_car_;
    std::vector<RobotRecord> assemble(const std::vector<Detection>& cars, const std::vector<std::vector<Detection>>& armors);
};

}  // namespace rmr
