#include "net.h"

#include <cstdlib>
#include <cstring>
#include <fstream>

namespace rmr {

namespace {
constexpr char kMagic[8] = {'R', 'M', 'R', 'E', 'N', 'G', '2', '\0'};
enum { OP_CONV = 0, OP_MAXPOOL5 = 1, OP_UPSAMPLE2 = 2, OP_COPY = 3 };

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
}

Net::~Net() {
    for (auto& kv : plans_)
        if (kv.second.graph) cudaGraphExecDestroy(kv.second.graph);
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
    return plans_.emplace(batch, std::move(bp)).first->second;
}

void Net::run_steps(const BatchPlan& bp, int batch, cudaStream_t s) {
    for (const Step& st : bp.steps) {
        const EngineOp& op = st.op;
        switch (st.type) {
            case OP_CONV:
                if (st.umma) launch_conv_umma(st.launch, s);
                else launch_conv_simt(st.desc, s);
                break;
            case OP_MAXPOOL5:
                launch_maxpool5(static_cast<const __half*>(bufs_[op.src_buf]), buf_desc_[op.src_buf].c, op.src_coff,
                                static_cast<__half*>(bufs_[op.dst_buf]), buf_desc_[op.dst_buf].c, op.dst_coff, batch,
                                op.src_h, op.src_w, op.src_c, s);
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
    if (bp.graph == nullptr) {
        conv_init();
        // capture on a private stream: the caller's stream may be the legacy default stream, which
        // cannot be captured, and nothing executes during capture anyway
        cudaStream_t cs = nullptr;
        RMR_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) {
            cudaStreamDestroy(cs);
            RMR_CUDA(e);
        }
        try {
            run_steps(bp, batch, cs);
        } catch (...) {
            cudaStreamEndCapture(cs, &g);
            if (g) cudaGraphDestroy(g);
            cudaStreamDestroy(cs);
            throw;
        }
        e = cudaStreamEndCapture(cs, &g);
        cudaStreamDestroy(cs);
        RMR_CUDA(e);
        RMR_CUDA(cudaGraphInstantiate(&bp.graph, g, 0));
        cudaGraphDestroy(g);
    }
    RMR_CUDA(cudaGraphLaunch(bp.graph, s));
}

}  // namespace rmr
