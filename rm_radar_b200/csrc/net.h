// Engine runtime: loads a `.rmeng` plan (rm_radar_b200/engine.py) and replays it as sm_100a launches.
// Stands where the reference holds a TensorRT engine + execution context
// (/root/reference/src/detect/detector.cpp:100-131, detector.h:122).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "conv.h"

namespace rmr {

struct EngineOp {
    int32_t type;
    int32_t src_buf, src_coff, src_c, src_h, src_w;
    int32_t dst_buf, dst_coff, dst_c, dst_h, dst_w;
    int32_t k, stride, act;
    int32_t res_buf, res_coff;
    int32_t cout_pad, cin_pad;
    int32_t reserved[6];
    int64_t w_off, b_off;
};
static_assert(sizeof(EngineOp) == 24 * 4 + 16, "EngineOp layout must match engine.py");

struct EngineBuf {
    int32_t h, w, c, dtype;   // dtype 0 = fp16, 1 = fp32
};

struct HeadLevel {
    const float* logits;      // [batch][h][w][pitch] fp32: 64 DFL bins then num_classes class logits
    int h, w, stride, pitch;
};

class Net {
public:
    Net(const std::string& engine_path, int max_batch);
    ~Net();
    Net(const Net&) = delete;
    Net& operator=(const Net&) = delete;

    // NHWC fp16 input, 4 channels per pixel (RGB + zero), [max_batch][in_h][in_w][4]
    __half* input() const { return static_cast<__half*>(bufs_[input_buf_]); }
    size_t input_stride() const { return static_cast<size_t>(in_h_) * in_w_ * 4; }
    void forward(int batch, cudaStream_t s);
    const std::vector<HeadLevel>& levels() const { return levels_; }
    int num_classes() const { return num_classes_; }
    int anchors() const { return anchors_; }
    int in_h() const { return in_h_; }
    int in_w() const { return in_w_; }
    int max_batch() const { return max_batch_; }
    double flops_per_image() const { return flops_per_image_; }
    int launches_per_forward() const { return static_cast<int>(ops_.size()); }
    // per-op timing (bench / profiling): ms per launch of every op at `batch`, `iters` back-to-back launches each
    std::vector<float> profile_ops(int batch, int iters, cudaStream_t s);
    void plan_stats(int batch, int* launches, int* umma_convs, int* lanes);
    bool op_uses_umma(int batch, int i) { return plan_for(batch).steps[i].umma; }
    // debugging / tests
    void set_force_simt(bool v) { force_simt_ = v; }
    void set_use_graph(bool v) { use_graph_ = v; }
    const void* buffer(int i) const { return bufs_[i]; }
    const EngineBuf& buffer_desc(int i) const { return buf_desc_[i]; }
    int num_buffers() const { return static_cast<int>(bufs_.size()); }
    const std::vector<EngineOp>& ops() const { return ops_; }

private:
    struct Step {
        int type;
        bool umma;
        ConvDesc desc;
        ConvLaunch launch;
        EngineOp op;
        int sppf_coff[3] = {0, 0, 0};   // OP_SPPF3: channel offsets of y1, y2, y3
        int lane = 0;               // capture stream this op is recorded on (independent branches overlap)
        std::vector<int> deps;      // ops on other lanes that must finish first (RAW / WAR / WAW on buffer views)
        bool signals = false;       // some op on another lane depends on this one
    };
    struct BatchPlan {
        std::vector<Step> steps;
        cudaGraphExec_t graph = nullptr;
        int lanes = 1;
        void* scratch = nullptr;   // split-K counters + partial tiles
    };
    BatchPlan& plan_for(int batch);
    void run_steps(const BatchPlan& bp, int batch, cudaStream_t s);
    void run_one(const Step& st, int batch, cudaStream_t s);
    void launch_step(const Step& st, int batch, cudaStream_t s, bool pdl);
    void schedule(BatchPlan& bp);
    void capture(BatchPlan& bp, int batch);

    int max_batch_ = 1, in_h_ = 0, in_w_ = 0, num_classes_ = 0, input_buf_ = 0, anchors_ = 0;
    std::vector<EngineBuf> buf_desc_;
    std::vector<void*> bufs_;
    std::vector<EngineOp> ops_;
    std::vector<HeadLevel> levels_;
    uint8_t* weights_ = nullptr;
    std::map<int, BatchPlan> plans_;
    bool force_simt_ = false;
    bool use_graph_ = true;
    int max_lanes_ = 8;
    double flops_per_image_ = 0;
};

}  // namespace rmr
