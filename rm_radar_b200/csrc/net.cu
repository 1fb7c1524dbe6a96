#include "net.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace rmr {

namespace {
constexpr char kMagic[8] = {'R', 'M', 'R', 'E', 'N', 'G', '2', '\0'};
enum { OP_CONV = 0, OP_MAXPOOL5 = 1, OP_UPSAMPLE2 = 2, OP_COPY = 3,
       OP_SPPF3 = 100, OP_FUSED_AWAY = 101 };   // runtime-only step types (never in the engine file)

struct Header {
    char magic[8];
    int32_t n_bufs, n_ops, n_levels, num_classes, in_h, in_w, input_buf, reserved;
    int64_t blob_bytes;
};
struct LevelRec {
    int32_t buf, h, w, stride;
};
}  // namespace

Net::Net(const std::string& engine_path, int max_batch) : max_batch_(max_batch) {
    std::ifstream f(engine_path, std::ios::binary);
    // same failure class as the reference's missing-engine path (detector.cpp:80: invalid_argument)
    if (!f) throw std::invalid_argument("engine file not found: " + engine_path +
                                        " (build it with `python -m rm_radar_b200.engine model.onnx model.rmeng`)");
    f.seekg(0, std::ios::end);
    const size_t size = static_cast<size_t>(f.tellg());
    f.seekg(0);
    std::vector<uint8_t> data(size);
    f.read(reinterpret_cast<char*>(data.data()), static_cast<std::streamsize>(size));
    if (size < sizeof(Header)) throw std::runtime_error("engine file truncated: " + engine_path);
    Header h;
    std::memcpy(&h, data.data(), sizeof(h));
    if (std::memcmp(h.magic, kMagic, 8) != 0) throw std::runtime_error("bad engine magic: " + engine_path);
    size_t pos = sizeof(Header);
    buf_desc_.resize(h.n_bufs);
    std::memcpy(buf_desc_.data(), data.data() + pos, sizeof(EngineBuf) * h.n_bufs);
    pos += sizeof(EngineBuf) * h.n_bufs;
    ops_.resize(h.n_ops);
    std::memcpy(ops_.data(), data.data() + pos, sizeof(EngineOp) * h.n_ops);
    pos += sizeof(EngineOp) * h.n_ops;
    std::vector<LevelRec> lv(h.n_levels);
    std::memcpy(lv.data(), data.data() + pos, sizeof(LevelRec) * h.n_levels);
    pos += sizeof(LevelRec) * h.n_levels;
    pos = (pos + 1023) / 1024 * 1024;
    if (pos + static_cast<size_t>(h.blob_bytes) > size) throw std::runtime_error("engine blob truncated");
    in_h_ = h.in_h; in_w_ = h.in_w; num_classes_ = h.num_classes; input_buf_ = h.input_buf;

    RMR_CUDA(cudaMalloc(&weights_, static_cast<size_t>(h.blob_bytes)));
    RMR_CUDA(cudaMemcpy(weights_, data.data() + pos, static_cast<size_t>(h.blob_bytes), cudaMemcpyHostToDevice));
    bufs_.resize(h.n_bufs, nullptr);
    for (int i = 0; i < h.n_bufs; ++i) {
        const EngineBuf& b = buf_desc_[i];
        const size_t bytes = static_cast<size_t>(max_batch_) * b.h * b.w * b.c * (b.dtype ? 4 : 2);
        RMR_CUDA(cudaMalloc(&bufs_[i], bytes));
        RMR_CUDA(cudaMemset(bufs_[i], 0, bytes));
    }
    for (const LevelRec& l : lv) {
        levels_.push_back(HeadLevel{static_cast<const float*>(bufs_[l.buf]), l.h, l.w, l.stride, buf_desc_[l.buf].c});
        anchors_ += l.h * l.w;
    }
    for (const EngineOp& op : ops_)
        if (op.type == OP_CONV)
            flops_per_image_ += 2.0 * op.dst_h * op.dst_w * static_cast<double>(op.dst_c) * op.k * op.k * op.src_c;
    const char* e = std::getenv("RMR_CONV_SIMT");
    force_simt_ = e && e[0] == '1';
    const char* g = std::getenv("RMR_NO_GRAPH");
    use_graph_ = !(g && g[0] == '1');
    const char* ln = std::getenv("RMR_GRAPH_LANES");
    if (ln) max_lanes_ = std::max(1, std::min(16, std::atoi(ln)));
}

Net::~Net() {
    for (auto& kv : plans_) {
        if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
        cudaFree(kv.second.scratch);
    }
    for (void* p : bufs_) cudaFree(p);
    cudaFree(weights_);
}

Net::BatchPlan& Net::plan_for(int batch) {
    auto it = plans_.find(batch);
    if (it != plans_.end()) return it->second;
    BatchPlan bp;
    for (const EngineOp& op : ops_) {
        Step st{};
        st.type = op.type;
        st.op = op;
        if (op.type == OP_CONV) {
            ConvDesc& d = st.desc;
            d.in = static_cast<const __half*>(bufs_[op.src_buf]);
            d.in_pitch = buf_desc_[op.src_buf].c; d.in_coff = op.src_coff; d.cin = op.src_c;
            d.h_in = op.src_h; d.w_in = op.src_w;
            d.out = bufs_[op.dst_buf];
            d.out_pitch = buf_desc_[op.dst_buf].c; d.out_coff = op.dst_coff; d.cout = op.dst_c;
            d.out_f32 = buf_desc_[op.dst_buf].dtype; d.h_out = op.dst_h; d.w_out = op.dst_w;
            d.k = op.k; d.stride = op.stride; d.act = op.act;
            if (op.res_buf >= 0) {
                d.res = static_cast<const __half*>(bufs_[op.res_buf]);
                d.res_pitch = buf_desc_[op.res_buf].c; d.res_coff = op.res_coff;
            }
            d.w = reinterpret_cast<const __half*>(weights_ + op.w_off);
            d.bias = reinterpret_cast<const float*>(weights_ + op.b_off);
            d.cout_pad = op.cout_pad; d.cin_pad = op.cin_pad;
            d.n = batch;
            st.umma = !force_simt_ && conv_umma_supported(d);
            if (st.umma) st.launch = make_conv_launch(d);
        }
        bp.steps.push_back(st);
    }
    // Upsample / Concat-copy of a conv result -> second destination of that conv's epilogue (no launch)
    for (size_t j = 0; j < bp.steps.size() && !force_simt_; ++j) {
        const EngineOp& u = bp.steps[j].op;
        if (bp.steps[j].type != OP_UPSAMPLE2 && bp.steps[j].type != OP_COPY) continue;
        for (size_t i = j; i-- > 0;) {
            Step& c = bp.steps[i];
            const EngineOp& co = c.op;
            const bool writes_src = co.dst_buf == u.src_buf && co.dst_coff < u.src_coff + u.src_c &&
                                    u.src_coff < co.dst_coff + (co.type == OP_CONV ? co.dst_c : co.src_c);
            if (!writes_src) continue;
            const bool exact = c.type == OP_CONV && c.umma && co.dst_coff == u.src_coff && co.dst_c == u.src_c &&
                               buf_desc_[co.dst_buf].dtype == 0 && c.desc.dup == nullptr && (co.dst_c % 8) == 0 &&
                               (buf_desc_[u.dst_buf].c % 8) == 0 && (u.dst_coff % 8) == 0 && c.launch.p.vec_ok;
            if (exact) {
                c.desc.dup = static_cast<__half*>(bufs_[u.dst_buf]);
                c.desc.dup_pitch = buf_desc_[u.dst_buf].c;
                c.desc.dup_coff = u.dst_coff;
                c.desc.dup_mode = bp.steps[j].type == OP_UPSAMPLE2 ? 2 : 1;
                c.launch = make_conv_launch(c.desc);
                bp.steps[j].type = OP_FUSED_AWAY;
            }
            break;   // the most recent writer of the source decides
        }
    }
    // split-K scratch: every split layer gets its own region (layers on different graph lanes overlap)
    size_t scratch = 0;
    for (const Step& st : bp.steps)
        if (st.umma) scratch += (conv_scratch_bytes(st.launch) + 255) / 256 * 256;
    if (scratch > 0) {
        RMR_CUDA(cudaMalloc(&bp.scratch, scratch));
        RMR_CUDA(cudaMemset(bp.scratch, 0, scratch));
        size_t off = 0;
        for (Step& st : bp.steps)
            if (st.umma) {
                const size_t b = (conv_scratch_bytes(st.launch) + 255) / 256 * 256;
                if (b) conv_bind_scratch(st.launch, static_cast<char*>(bp.scratch) + off);
                off += b;
            }
    }
    // SPPF: three chained 5x5 max-pools writing into one buffer -> a single launch
    for (size_t i = 0; i + 2 < bp.steps.size(); ++i) {
        const EngineOp &a = bp.steps[i].op, &b = bp.steps[i + 1].op, &c = bp.steps[i + 2].op;
        if (a.type != OP_MAXPOOL5 || b.type != OP_MAXPOOL5 || c.type != OP_MAXPOOL5) continue;
        const bool chained = b.src_buf == a.dst_buf && b.src_coff == a.dst_coff && c.src_buf == b.dst_buf &&
                             c.src_coff == b.dst_coff && a.dst_buf == b.dst_buf && b.dst_buf == c.dst_buf &&
                             a.src_c == b.src_c && b.src_c == c.src_c && a.src_h * a.src_w <= 1024 &&
                             (a.src_c % 8) == 0;
        if (!chained || force_simt_) continue;
        bp.steps[i].type = OP_SPPF3;
        bp.steps[i + 1].type = OP_FUSED_AWAY;
        bp.steps[i + 2].type = OP_FUSED_AWAY;
        bp.steps[i].sppf_coff[0] = a.dst_coff; bp.steps[i].sppf_coff[1] = b.dst_coff; bp.steps[i].sppf_coff[2] = c.dst_coff;
    }
    schedule(bp);
    return plans_.emplace(batch, std::move(bp)).first->second;
}

void Net::launch_step(const Step& st, int batch, cudaStream_t s, bool pdl) {
    const EngineOp& op = st.op;
    switch (st.type) {
        case OP_CONV:
            if (st.umma) launch_conv_umma(st.launch, s, pdl);
            else launch_conv_simt(st.desc, s);
            break;
        case OP_MAXPOOL5:
            launch_maxpool5(static_cast<const __half*>(bufs_[op.src_buf]), buf_desc_[op.src_buf].c, op.src_coff,
                            static_cast<__half*>(bufs_[op.dst_buf]), buf_desc_[op.dst_buf].c, op.dst_coff, batch,
                            op.src_h, op.src_w, op.src_c, s);
            break;
        case OP_SPPF3:
            launch_sppf_pool3(static_cast<const __half*>(bufs_[op.src_buf]), buf_desc_[op.src_buf].c, op.src_coff,
                              static_cast<__half*>(bufs_[op.dst_buf]), buf_desc_[op.dst_buf].c, st.sppf_coff[0],
                              st.sppf_coff[1], st.sppf_coff[2], batch, op.src_h, op.src_w, op.src_c, s);
            break;
        case OP_FUSED_AWAY:
            break;
        case OP_UPSAMPLE2:
            launch_upsample2(static_cast<const __half*>(bufs_[op.src_buf]), buf_desc_[op.src_buf].c, op.src_coff,
                             static_cast<__half*>(bufs_[op.dst_buf]), buf_desc_[op.dst_buf].c, op.dst_coff, batch,
                             op.src_h, op.src_w, op.src_c, s);
            break;
        case OP_COPY:
            launch_copy_channels(static_cast<const __half*>(bufs_[op.src_buf]), buf_desc_[op.src_buf].c,
                                 op.src_coff, static_cast<__half*>(bufs_[op.dst_buf]), buf_desc_[op.dst_buf].c,
                                 op.dst_coff, batch, op.src_h, op.src_w, op.src_c, s);
            break;
        default:
            throw std::runtime_error("unknown engine op");
    }
}

void Net::run_steps(const BatchPlan& bp, int batch, cudaStream_t s) {
    for (const Step& st : bp.steps) launch_step(st, batch, s, true);
}

// Dependency analysis over buffer views (buffer, channel range): RAW, WAR and WAW hazards give the
// partial order; ops are then laid out on up to max_lanes_ capture streams so that independent
// branches (the 2 x 4 head branches, head levels vs. the rest of the neck) become parallel graph
// branches.  An op continues the lane of its most recent producer when that producer is still the
// lane's tail; otherwise it opens / reuses another lane.
void Net::schedule(BatchPlan& bp) {
    struct View { int buf, c0, c1; };
    const int n = static_cast<int>(bp.steps.size());
    auto overlap = [](const View& a, const View& b) { return a.buf == b.buf && a.c0 < b.c1 && b.c0 < a.c1; };
    std::vector<std::vector<View>> reads(n), writes(n);
    for (int i = 0; i < n; ++i) {
        const EngineOp& op = bp.steps[i].op;
        reads[i].push_back(View{op.src_buf, op.src_coff, op.src_coff + op.src_c});
        if (op.type == OP_CONV && op.res_buf >= 0) reads[i].push_back(View{op.res_buf, op.res_coff, op.res_coff + op.dst_c});
        const int wc = (op.type == OP_CONV) ? op.dst_c : op.src_c;
        writes[i].push_back(View{op.dst_buf, op.dst_coff, op.dst_coff + wc});
    }
    std::vector<std::vector<int>> preds(n);
    for (int i = 0; i < n; ++i)
        for (int j = i - 1; j >= 0; --j) {
            bool dep = false;
            for (const View& w : writes[j]) {
                for (const View& r : reads[i]) dep |= overlap(w, r);    // RAW
                for (const View& w2 : writes[i]) dep |= overlap(w, w2); // WAW
            }
            for (const View& r : reads[j])
                for (const View& w2 : writes[i]) dep |= overlap(r, w2); // WAR
            if (dep) preds[i].push_back(j);
        }
    std::vector<int> tail(max_lanes_, -1);      // last op recorded on each lane
    int lanes_used = 1;
    for (int i = 0; i < n; ++i) {
        Step& st = bp.steps[i];
        int lane = -1;
        for (int j : preds[i])                   // preds are in descending order: most recent producer first
            if (tail[bp.steps[j].lane] == j) { lane = bp.steps[j].lane; break; }
        if (lane < 0) {
            if (i == 0) lane = 0;
            else if (lanes_used < max_lanes_) lane = lanes_used++;
            else {
                // all lanes busy: queue behind the lane whose tail is oldest
                lane = 0;
                for (int l = 1; l < max_lanes_; ++l)
                    if (tail[l] < tail[lane]) lane = l;
            }
        }
        st.lane = lane;
        st.deps.clear();
        for (int j : preds[i])
            if (bp.steps[j].lane != lane) {
                st.deps.push_back(j);
                bp.steps[j].signals = true;
            }
        tail[lane] = i;
    }
    bp.lanes = lanes_used;
}

// Stream capture of the scheduled plan: lane 0 is the origin stream, other lanes fork from it through
// event waits and are joined back before EndCapture.  Consecutive tcgen05 convs on one lane keep their
// programmatic (PDL) edge; an op that also waits on another lane is launched with a full dependency.
void Net::capture(BatchPlan& bp, int batch) {
    conv_init();
    const int n = static_cast<int>(bp.steps.size());
    std::vector<cudaStream_t> ls(bp.lanes, nullptr);
    std::vector<cudaEvent_t> ev(n, nullptr);
    std::vector<cudaEvent_t> extra;
    cudaGraph_t g = nullptr;
    auto cleanup = [&] {
        for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : extra) cudaEventDestroy(e);
        for (cudaStream_t s : ls) if (s) cudaStreamDestroy(s);
    };
    try {
        for (auto& s : ls) RMR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        RMR_CUDA(cudaStreamBeginCapture(ls[0], cudaStreamCaptureModeThreadLocal));
        cudaEvent_t fork;
        RMR_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        extra.push_back(fork);
        RMR_CUDA(cudaEventRecord(fork, ls[0]));
        std::vector<bool> joined(bp.lanes, false);
        joined[0] = true;
        for (int i = 0; i < n; ++i) {
            const Step& st = bp.steps[i];
            cudaStream_t s = ls[st.lane];
            if (!joined[st.lane]) {
                RMR_CUDA(cudaStreamWaitEvent(s, fork, 0));
                joined[st.lane] = true;
            }
            for (int j : st.deps) RMR_CUDA(cudaStreamWaitEvent(s, ev[j], 0));
            launch_step(st, batch, s, st.deps.empty());
            if (st.signals) {
                RMR_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
                RMR_CUDA(cudaEventRecord(ev[i], s));
            }
        }
        for (int l = 1; l < bp.lanes; ++l) {
            if (!joined[l]) continue;
            cudaEvent_t e;
            RMR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            extra.push_back(e);
            RMR_CUDA(cudaEventRecord(e, ls[l]));
            RMR_CUDA(cudaStreamWaitEvent(ls[0], e, 0));
        }
        RMR_CUDA(cudaStreamEndCapture(ls[0], &g));
        RMR_CUDA(cudaGraphInstantiate(&bp.graph, g, 0));
    } catch (...) {
        cudaGraph_t dead = nullptr;
        cudaStreamCaptureStatus cst;
        if (ls[0] && cudaStreamIsCapturing(ls[0], &cst) == cudaSuccess && cst != cudaStreamCaptureStatusNone)
            cudaStreamEndCapture(ls[0], &dead);
        if (dead) cudaGraphDestroy(dead);
        if (g) cudaGraphDestroy(g);
        cleanup();
        throw;
    }
    cudaGraphDestroy(g);
    cleanup();
}

void Net::plan_stats(int batch, int* launches, int* umma_convs, int* lanes) {
    if (batch <= 0 || batch > max_batch_) throw std::invalid_argument("plan_stats: bad batch");
    const BatchPlan& bp = plan_for(batch);
    int nl = 0, nu = 0;
    for (const Step& st : bp.steps) {
        if (st.type == OP_FUSED_AWAY) continue;
        ++nl;
        if (st.type == OP_CONV && st.umma) ++nu;
    }
    if (launches) *launches = nl;
    if (umma_convs) *umma_convs = nu;
    if (lanes) *lanes = bp.lanes;
}

void Net::run_one(const Step& st, int batch, cudaStream_t s) {
    BatchPlan one;
    one.steps.push_back(st);
    run_steps(one, batch, s);
}

std::vector<float> Net::profile_ops(int batch, int iters, cudaStream_t s) {
    if (batch <= 0 || batch > max_batch_ || iters <= 0) throw std::invalid_argument("profile_ops: bad batch / iters");
    conv_init();
    BatchPlan& bp = plan_for(batch);
    std::vector<float> ms(bp.steps.size(), 0.f);
    cudaEvent_t e0, e1;
    RMR_CUDA(cudaEventCreate(&e0));
    RMR_CUDA(cudaEventCreate(&e1));
    for (size_t i = 0; i < bp.steps.size(); ++i) {
        run_one(bp.steps[i], batch, s);   // warm
        RMR_CUDA(cudaEventRecord(e0, s));
        for (int k = 0; k < iters; ++k) run_one(bp.steps[i], batch, s);
        RMR_CUDA(cudaEventRecord(e1, s));
        RMR_CUDA(cudaEventSynchronize(e1));
        RMR_CUDA(cudaEventElapsedTime(&ms[i], e0, e1));
        ms[i] /= static_cast<float>(iters);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

void Net::forward(int batch, cudaStream_t s) {
    if (batch <= 0) return;
    if (batch > max_batch_) throw std::invalid_argument("batch exceeds max_batch");
    BatchPlan& bp = plan_for(batch);
    if (!use_graph_) {
        run_steps(bp, batch, s);
        return;
    }
    if (bp.graph == nullptr) capture(bp, batch);
    RMR_CUDA(cudaGraphLaunch(bp.graph, s));
}

}  // namespace rmr
