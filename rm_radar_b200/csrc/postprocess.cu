#include "postprocess.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

namespace rmr {

namespace {

struct LevelArg {
    const float* logits;
    int h, w, stride, pitch, anchor0;
};
struct LevelsArg {
    LevelArg lv[4];
    int n_levels, anchors;
};

// Per anchor: class sigmoid + first-max argmax (decodeKernel, detector.cu:229-235), confidence
// threshold (NMSKernel's `row_conf < score_thresh`, detector.cu:339-343, hoisted in front so that
// only survivors are decoded), then the exported Detect tail for that anchor (DFL softmax over 16
// bins, dist2bbox, x stride — SURVEY.md Appendix A) and decodeKernel's cxcywh -> clamped xywh.
__device__ __forceinline__ void decode_anchor(const LevelsArg& la, int num_classes, float conf_thresh, float* __restrict__ cand,
                                              int* __restrict__ cand_count) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (a >= la.anchors) return;
    int l = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < la.n_levels && a >= la.lv[i].anchor0) l = i;
    const LevelArg lv = la.lv[l];
    const int local = a - lv.anchor0;
    const int px = local % lv.w, py = local / lv.w;
    const float* row = lv.logits + (static_cast<size_t>(img) * lv.h * lv.w + local) * lv.pitch;

    float best = -1.f;
    int label = 0;
    for (int c = 0; c < num_classes; ++c) {
        const float s = 1.f / (1.f + expf(-row[64 + c]));
        if (s > best) { best = s; label = c; }
    }
    if (best < conf_thresh) return;

    float dist[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(row + side * 16 + j * 4);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
        float m = v[0];
#pragma unroll
        for (int j = 1; j < 16; ++j) m = fmaxf(m, v[j]);
        float sum = 0.f, acc = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float e = expf(v[j] - m);
            sum += e;
            acc += e * static_cast<float>(j);
        }
        dist[side] = acc / sum;
    }
    const float ax = static_cast<float>(px) + 0.5f, ay = static_cast<float>(py) + 0.5f;
    const float x1 = ax - dist[0], y1 = ay - dist[1], x2 = ax + dist[2], y2 = ay + dist[3];
    const float st = static_cast<float>(lv.stride);
    const float cx = (x1 + x2) * 0.5f * st, cy = (y1 + y2) * 0.5f * st;
    const float w = (x2 - x1) * st, h = (y2 - y1) * st;
    // decodeKernel: the `0.5 *` is a double literal (detector.cu:237-238)
    const float x = static_cast<float>(fmax(static_cast<double>(cx) - 0.5 * static_cast<double>(w), 0.0));
    const float y = static_cast<float>(fmax(static_cast<double>(cy) - 0.5 * static_cast<double>(h), 0.0));

    const int slot = atomicAdd(cand_count + img, 1);
    if (slot >= kMaxCandidates) return;
    float* o = cand + (static_cast<size_t>(img) * kMaxCandidates + slot) * 8;
    reinterpret_cast<float4*>(o)[0] = make_float4(x, y, w, h);
    reinterpret_cast<float4*>(o)[1] = make_float4(static_cast<float>(label), best, __int_as_float(a), 0.f);
}

// IoU — detector.cu:271-293
__device__ __forceinline__ float iou_xywh(const float4 a, const float4 b) {
    const float xl = fmaxf(a.x, b.x), yt = fmaxf(a.y, b.y);
    const float xr = fminf(a.x + a.z, b.x + b.z), yb = fminf(a.y + a.w, b.y + b.w);
    if (xr < xl || yb < yt) return 0.f;
    const float inter = (xr - xl) * (yb - yt);
    const float uni = a.z * a.w + b.z * b.w - inter;
    return inter / uni;
}

// One block per image.  Candidates arrive in arbitrary (atomic) order; they are first ranked by
// anchor index so that the output order is the reference's anchor order (detector.cu:561-579),
// then every row is tested against every column (race-free all-pairs rule, Appendix B#5/#6),
// survivors are compacted in order and un-letterboxed (restoreDetection, detector.cpp:258-268).
// Up to kSmemCand candidates (every real frame) the ranked list lives in shared memory, so the
// two all-pairs loops read broadcast words instead of L1 lines; longer lists use the global scratch.
constexpr int kSmemCand = 1024;

template <bool kShared>
__device__ __forceinline__ void nms_restore_body(const float* c, int n, float4* sbox, float4* smeta,
                                                 int* sanchor, const LetterboxGeom& g, float nms_thresh,
                                                 Detection* __restrict__ out, int* __restrict__ out_count, int max_out,
                                                 int* warp_tot, int* base, Detection* __restrict__ host_out, int host_head,
                                                 int* __restrict__ host_out_count) {
    if (kShared) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sanchor[i] = __float_as_int(__ldcg(c + i * 8 + 6));
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float4 b0 = __ldcg(reinterpret_cast<const float4*>(c + i * 8));
        const float4 b1 = __ldcg(reinterpret_cast<const float4*>(c + i * 8) + 1);
        const int anchor = __float_as_int(b1.z);
        int rank = 0;
        if (kShared) {
            for (int j = 0; j < n; ++j) rank += (sanchor[j] < anchor) ? 1 : 0;
        } else {
            for (int j = 0; j < n; ++j) rank += (__float_as_int(__ldcg(c + j * 8 + 6)) < anchor) ? 1 : 0;
        }
        sbox[kShared ? rank : 2 * rank] = b0;
        smeta[kShared ? rank : 2 * rank] = b1;
    }
    if (threadIdx.x == 0) *base = 0;
    __syncthreads();

    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool keep = false;
        float4 box = make_float4(0, 0, 0, 0);
        float label = 0.f, conf = 0.f;
        if (i < n) {
            box = sbox[kShared ? i : 2 * i];
            const float4 m = smeta[kShared ? i : 2 * i];
            label = m.x; conf = m.y;
            keep = true;
            for (int j = 0; j < n && keep; ++j) {
                const float4 mj = smeta[kShared ? j : 2 * j];
                if (mj.x == label && mj.y > conf) {
                    const float4 bj = sbox[kShared ? j : 2 * j];
                    if (iou_xywh(box, bj) > nms_thresh) keep = false;
                }
            }
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warp_tot[warp] = __popc(ballot);
        __syncthreads();
        int off = *base;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        off += __popc(ballot & ((1u << lane) - 1u));
        if (keep && off < max_out) {
            Detection d;
            d.x = fminf(fmaxf((box.x - g.dw) * g.ratio, 0.f), g.width);
            d.y = fminf(fmaxf((box.y - g.dh) * g.ratio, 0.f), g.height);
            d.width = fminf(fmaxf(box.z * g.ratio, 0.f), g.width - d.x);
            d.height = fminf(fmaxf(box.w * g.ratio, 0.f), g.height - d.y);
            d.label = label;
            d.confidence = conf;
            out[off] = d;
            if (off < host_head) host_out[off] = d;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_tot[w];
            *base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out_count = *base;
        if (host_out_count) *host_out_count = *base;
    }
}

struct NmsArgs {
    const float* cand;
    const int* cand_count;
    float* sorted_scratch;
    const LetterboxGeom* geoms;
    float nms_thresh;
    Detection* out;
    int* out_count;
    int max_out;
    Detection* host_out;
    int host_head;
    int* host_out_count;
    int* host_cand_count;
    int* host_done;
    int seq;
};

struct NmsShared {
    float4 box[kSmemCand];
    float4 meta[kSmemCand];
    int anchor[kSmemCand];
    int warp_tot[8];
    int base;
};

__device__ __forceinline__ void nms_restore_image(const NmsArgs& a, int img, NmsShared& sh) {
    const int n_raw = __ldcg(a.cand_count + img);
    if (a.host_cand_count && threadIdx.x == 0) a.host_cand_count[img] = n_raw;
    const int n = min(n_raw, kMaxCandidates);
    const float* c = a.cand + static_cast<size_t>(img) * kMaxCandidates * 8;
    const LetterboxGeom g = a.geoms[img];
    Detection* o = a.out + static_cast<size_t>(img) * a.max_out;
    Detection* ho = a.host_out ? a.host_out + static_cast<size_t>(img) * a.max_out : nullptr;
    const int head = a.host_out ? a.host_head : 0;
    int* hc = a.host_out_count ? a.host_out_count + img : nullptr;
    if (n <= kSmemCand) {
        nms_restore_body<true>(c, n, sh.box, sh.meta, sh.anchor, g, a.nms_thresh, o, a.out_count + img, a.max_out, sh.warp_tot,
                               &sh.base, ho, head, hc);
    } else {
        // global scratch rows are [box float4][meta float4]: element i of either view sits at float4 index 2 i
        float4* sorted = reinterpret_cast<float4*>(a.sorted_scratch + static_cast<size_t>(img) * kMaxCandidates * 8);
        nms_restore_body<false>(c, n, sorted, sorted + 1, nullptr, g, a.nms_thresh, o, a.out_count + img, a.max_out, sh.warp_tot,
                                &sh.base, ho, head, hc);
    }
}

// the NMS stage alone (self-test hook)
__global__ void __launch_bounds__(256) nms_restore_kernel(const NmsArgs a) {
    __shared__ NmsShared sh;
    nms_restore_image(a, blockIdx.x, sh);
    if (a.host_done) {
        // every thread's host-visible writes are ordered before the flag: fence, block barrier, one store
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(a.host_done + blockIdx.x) = a.seq;
    }
}

__global__ void __launch_bounds__(256) decode_compact_kernel(LevelsArg la, int num_classes, float conf_thresh, float* cand,
                                                             int* cand_count) {
    decode_anchor(la, num_classes, conf_thresh, cand, cand_count);
}

}  // namespace

void post_alloc(PostBuffers& pb, int max_batch, int max_out) {
    pb.max_batch = max_batch;
    pb.max_out = max_out;
    // cand holds two regions: raw candidates and the anchor-sorted copy
    RMR_CUDA(cudaMalloc(&pb.cand, sizeof(float) * 8 * kMaxCandidates * max_batch * 2));
    RMR_CUDA(cudaMalloc(&pb.cand_count, sizeof(int) * max_batch));
    RMR_CUDA(cudaMalloc(&pb.out, sizeof(Detection) * max_out * max_batch));
    RMR_CUDA(cudaMalloc(&pb.out_count, sizeof(int) * max_batch));
}

void post_free(PostBuffers& pb) {
    cudaFree(pb.cand); cudaFree(pb.cand_count); cudaFree(pb.out); cudaFree(pb.out_count);
    pb = PostBuffers{};
}

void launch_postprocess(const std::vector<HeadLevel>& levels, int num_classes, int batch,
                        const LetterboxGeom* dev_geoms, float conf_thresh, float nms_thresh, PostBuffers& pb,
                        cudaStream_t s) {
    if (batch <= 0) return;
    LevelsArg la{};
    la.n_levels = static_cast<int>(levels.size());
    int a0 = 0;
    for (int i = 0; i < la.n_levels; ++i) {
        la.lv[i] = LevelArg{levels[i].logits, levels[i].h, levels[i].w, levels[i].stride, levels[i].pitch, a0};
        a0 += levels[i].h * levels[i].w;
    }
    la.anchors = a0;
    RMR_CUDA(cudaMemsetAsync(pb.cand_count, 0, sizeof(int) * batch, s));
    float* sorted = pb.cand + static_cast<size_t>(8) * kMaxCandidates * pb.max_batch;
    const NmsArgs na{pb.cand, pb.cand_count, sorted, dev_geoms, nms_thresh, pb.out, pb.out_count, pb.max_out,
                     pb.host_out, pb.host_head, pb.host_out_count, pb.host_cand_count, const_cast<int*>(pb.host_done), pb.seq};
    // (decode + NMS as one launch — the last decode block of an image running its NMS — measured no faster than two
    // launches: 1.3033 vs 1.3025 ms per step)
    decode_compact_kernel<<<dim3((a0 + 255) / 256, batch), 256, 0, s>>>(la, num_classes, conf_thresh, pb.cand, pb.cand_count);
    nms_restore_kernel<<<batch, 256, 0, s>>>(na);
    RMR_CUDA(cudaGetLastError());
}

std::vector<Detection> postprocess_selftest(const float* cand6, int n, float nms_thresh) {
    if (n < 0) throw std::invalid_argument("negative candidate count");
    if (n > kMaxCandidates)
        throw CapacityError(std::to_string(n) + " anchors pass the confidence threshold, capacity " + std::to_string(kMaxCandidates));
    constexpr int kOut = 4096;
    PostBuffers pb;
    post_alloc(pb, 1, kOut);
    std::vector<float> h(static_cast<size_t>(std::max(n, 1)) * 8, 0.f);
    for (int i = 0; i < n; ++i) {            // reversed arrival order, anchor = original row
        float* o = h.data() + static_cast<size_t>(n - 1 - i) * 8;
        for (int k = 0; k < 6; ++k) o[k] = cand6[i * 6 + k];
        std::memcpy(o + 6, &i, sizeof(int));
    }
    LetterboxGeom g{};
    g.ratio = 1.f; g.dw = 0.f; g.dh = 0.f; g.width = 1e9f; g.height = 1e9f;
    LetterboxGeom* dg = nullptr;
    std::vector<Detection> out;
    try {
        RMR_CUDA(cudaMalloc(&dg, sizeof(g)));
        RMR_CUDA(cudaMemcpy(dg, &g, sizeof(g), cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemcpy(pb.cand, h.data(), sizeof(float) * 8 * n, cudaMemcpyHostToDevice));
        RMR_CUDA(cudaMemcpy(pb.cand_count, &n, sizeof(int), cudaMemcpyHostToDevice));
        float* sorted = pb.cand + static_cast<size_t>(8) * kMaxCandidates * pb.max_batch;
        nms_restore_kernel<<<1, 256>>>(NmsArgs{pb.cand, pb.cand_count, sorted, dg, nms_thresh, pb.out, pb.out_count, pb.max_out,
                                               nullptr, 0, nullptr, nullptr, nullptr, 0});
        RMR_CUDA(cudaGetLastError());
        int cnt = 0;
        RMR_CUDA(cudaMemcpy(&cnt, pb.out_count, sizeof(int), cudaMemcpyDeviceToHost));
        if (cnt > kOut) throw CapacityError(std::to_string(cnt) + " detections survive NMS, self-test capacity " + std::to_string(kOut));
        out.resize(cnt);
        if (cnt) RMR_CUDA(cudaMemcpy(out.data(), pb.out, sizeof(Detection) * cnt, cudaMemcpyDeviceToHost));
    } catch (...) {
        cudaFree(dg);
        post_free(pb);
        throw;
    }
    cudaFree(dg);
    post_free(pb);
    return out;
}

}  // namespace rmr
