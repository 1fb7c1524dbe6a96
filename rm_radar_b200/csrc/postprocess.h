#pragma once
#include <vector>

#include "common.cuh"
#include "net.h"
#include "preprocess.h"

namespace rmr {

// Same POD as radar::Detection (/root/reference/src/detect/detection.h:25-68): six floats.
struct Detection {
    float x, y, width, height, label, confidence;
};
static_assert(sizeof(Detection) == 24, "Detection must stay a 6-float POD");

constexpr int kMaxCandidates = 16384;   // per image, after the confidence threshold; exceeding it is a CapacityError

struct PostBuffers {
    // device
    float* cand = nullptr;        // [batch][kMaxCandidates][8]: x,y,w,h,label,conf,anchor,pad
    int* cand_count = nullptr;    // [batch]
    Detection* out = nullptr;     // [batch][max_out]
    int* out_count = nullptr;     // [batch] (survivors, may exceed max_out -> truncated on write)
    int max_out = 0;
    int max_batch = 0;
    // optional host mirrors (pinned, device-visible): the NMS kernel writes the two counters and the first host_head
    // detections of every image straight into them, so the step needs no device->host copy (three small copies cost
    // 14-20 us per stage, profiles/r2_summary.md); longer lists are fetched from `out` on demand
    Detection* host_out = nullptr;     // [batch][max_out], first host_head rows of each image are written
    int* host_out_count = nullptr;     // [batch]
    int* host_cand_count = nullptr;    // [batch]
    int host_head = 0;
    // completion flags: block `img` of the NMS kernel stores `seq` into host_done[img] after a system-wide fence, so the
    // host can watch the pinned word instead of paying a stream-synchronise wake-up
    volatile int* host_done = nullptr; // [batch]
    int seq = 0;
};

void post_alloc(PostBuffers& pb, int max_batch, int max_out);
void post_free(PostBuffers& pb);

// Fused replacement for transposeKernel + decodeKernel + NMSKernel + the CPU NaN filter and
// restoreDetection (detector.cu:185-360, 522-582; detector.cpp:258-268): reads the head logits in
// place, thresholds first, and only the survivors reach the all-pairs NMS.
void launch_postprocess(const std::vector<HeadLevel>& levels, int num_classes, int batch,
                        const LetterboxGeom* dev_geoms, float conf_thresh, float nms_thresh, PostBuffers& pb,
                        cudaStream_t s);

// Test hook (tests only): nms_restore_kernel on caller-supplied candidates [n][6] = (x, y, w, h, label, conf) in
// network coordinates, row index = anchor index, uploaded in reverse order (the kernel must restore anchor order);
// identity un-letterboxing.  More than kMaxCandidates rows is a CapacityError, like in the product path.
std::vector<Detection> postprocess_selftest(const float* cand6, int n, float nms_thresh);

}  // namespace rmr
