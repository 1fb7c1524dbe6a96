#pragma once
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace rmr {

// Letterbox geometry of one image / ROI: PreParam (preparam.h:25-58) + the integer kernel
// arguments the reference derives from it at the call sites (detector.cu:393-410).
struct LetterboxGeom {
    int src_x, src_y, src_w, src_h;   // source rectangle inside the frame
    int pw, ph;                       // resized size (int-truncated in compat mode)
    int top, left;                    // border offsets
    int bw, bh;                       // bordered size actually written (<= 640)
    int stride_w;                     // row stride (pixels) the border kernel writes with
    int clean;                        // 1: bordered == out_w x out_h -> single fused kernel
    float width, height, ratio, dw, dh;   // PreParam, used by restoreDetection
};

LetterboxGeom make_letterbox_geom(int src_x, int src_y, int src_w, int src_h, int out_w, int out_h, bool compat);

// `frame`: device BGR u8, `stride` bytes per row.  `out`: [count][out_h][out_w][4] fp16 (R,G,B,0).
// `staging`: [count][out_h*out_w*3] persistent u8 (the reference's dev_border_ptr_).
void launch_letterbox(const unsigned char* frame, int stride, const LetterboxGeom* dev_geoms, bool any_unclean,
                      bool any_clean, int count, unsigned char* staging, __half* out, int out_w, int out_h,
                      cudaStream_t s);
// kernels the compat (staged) path launches per call: 1 fused, 2 with RMR_LETTERBOX_TWO_PASS=1
int letterbox_compat_launches();

}  // namespace rmr
