// Engine builder in the library: ONNX file -> `.rmeng` plan, and the reference's engine-path resolution.
//
// The reference's Detector is handed `<x>.engine`; when that cache is absent it parses the sibling `<x>.onnx` with
// nvonnxparser, builds an FP16 TensorRT engine and writes it back (/root/reference/src/detect/detector.cpp:74-99,
// 177-243, 281-311).  Here the plan is the flat file net.cu replays (header | buffers | ops | levels | weight blob),
// and the builder does at build time what TensorRT's fusion passes do: Conv+bias+SiLU (+ shortcut Add) become one
// conv op, Concat / Split / Slice become channel-offset views of one buffer, the exported Detect tail becomes the
// DECODE stage of postprocess.cu.  Host code only; the file it writes is byte-identical to the one
// rm_radar_b200/engine.py writes (tests/test_engine_cc.py), so either builder serves either runtime.
#include "engine.h"

#include <sys/stat.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <stdexcept>
#include <unordered_map>

#include "net.h"

namespace rmr {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// protobuf wire format: (field << 3 | type) varint keys; type 0 varint, 1 fixed64, 2 length-delimited, 5 fixed32.
// Only the ONNX fields the two graphs use are read.
// ------------------------------------------------------------------------------------------------------------------
struct Span {
    const uint8_t* p = nullptr;
    size_t n = 0;
    std::string str() const { return std::string(reinterpret_cast<const char*>(p), n); }
};

class Fields {
public:
    explicit Fields(Span s) : cur_(s.p), end_(s.p + s.n) {}
    // next record of the message; false at the end
    bool next() {
        if (cur_ >= end_) return false;
        const uint64_t key = varint();
        field = static_cast<uint32_t>(key >> 3);
        wire = static_cast<uint32_t>(key & 7);
        switch (wire) {
            case 0: value = varint(); break;
            case 1: need(8); std::memcpy(&value, cur_, 8); cur_ += 8; break;
            case 2: {
                const uint64_t len = varint();
                need(len);
                bytes = Span{cur_, static_cast<size_t>(len)};
                cur_ += len;
                break;
            }
            case 5: { need(4); uint32_t v; std::memcpy(&v, cur_, 4); value = v; cur_ += 4; break; }
            default: throw std::runtime_error("onnx: unsupported protobuf wire type " + std::to_string(wire));
        }
        return true;
    }
    uint32_t field = 0, wire = 0;
    uint64_t value = 0;
    Span bytes;

private:
    void need(uint64_t n) const {
        if (static_cast<uint64_t>(end_ - cur_) < n) throw std::runtime_error("onnx: truncated protobuf record");
    }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 70; shift += 7) {
            need(1);
            const uint8_t b = *cur_++;
            v |= static_cast<uint64_t>(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        throw std::runtime_error("onnx: varint too long");
    }
    const uint8_t *cur_, *end_;
};

void packed_ints(Span s, std::vector<int64_t>& out) {
    const uint8_t *p = s.p, *e = s.p + s.n;
    while (p < e) {
        uint64_t v = 0;
        int shift = 0;
        for (;;) {
            if (p >= e) throw std::runtime_error("onnx: truncated packed varint");
            const uint8_t b = *p++;
            v |= static_cast<uint64_t>(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
            shift += 7;
        }
        out.push_back(static_cast<int64_t>(v));
    }
}

struct Tensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 0;
    std::vector<float> f;      // float tensors (fp32 / fp16 / fp64 sources), row-major
    std::vector<int64_t> i;    // integer tensors
    int64_t first_int() const {
        if (i.empty()) throw std::runtime_error("onnx: integer initializer expected: " + name);
        return i[0];
    }
};

float half_bits_to_float(uint16_t h) {
    const uint32_t sign = static_cast<uint32_t>(h & 0x8000) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { man <<= 1; ++e; } while (!(man & 0x400));
            bits = sign | static_cast<uint32_t>(127 - 15 - e) << 23 | (man & 0x3ff) << 13;
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
    else bits = sign | (exp + 112) << 23 | man << 13;
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

// fp32 -> fp16 bits, round to nearest even, overflow to infinity (what numpy's astype(float16) does)
uint16_t float_to_half_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint16_t sign = static_cast<uint16_t>((x >> 16) & 0x8000);
    const uint32_t abs = x & 0x7fffffffu;
    if (abs >= 0x7f800000u) return sign | 0x7c00 | (abs > 0x7f800000u ? 0x200 | ((abs >> 13) & 0x3ff) : 0);
    if (abs >= 0x477ff000u) return sign | 0x7c00;                      // rounds to >= 65520 -> inf
    if (abs < 0x33000001u) return sign;                                 // <= 2^-25 rounds to zero (tie goes to even = 0)
    const int exp = static_cast<int>(abs >> 23) - 127;
    uint32_t man = (abs & 0x7fffffu) | 0x800000u;
    int shift;
    uint32_t base;
    if (exp < -14) { shift = 13 + (-14 - exp); base = 0; }              // subnormal half
    else { shift = 13; base = static_cast<uint32_t>(exp + 15) << 10; man &= 0x7fffffu; }
    uint32_t q = man >> shift;
    const uint32_t rem = man & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) ++q;                  // carries propagate into the exponent field
    return sign | static_cast<uint16_t>(base + q);
}

Tensor parse_tensor(Span s) {
    Tensor t;
    Span raw;
    bool has_raw = false;
    std::vector<float> floats;
    std::vector<int64_t> ints;
    for (Fields f(s); f.next();) {
        switch (f.field) {
            case 1: if (f.wire == 2) packed_ints(f.bytes, t.dims); else t.dims.push_back(static_cast<int64_t>(f.value)); break;
            case 2: t.dtype = static_cast<int>(f.value); break;
            case 4:
                if (f.wire == 2) {
                    const size_t k = f.bytes.n / 4, at = floats.size();
                    floats.resize(at + k);
                    std::memcpy(floats.data() + at, f.bytes.p, k * 4);
                } else { const uint32_t v = static_cast<uint32_t>(f.value); float x; std::memcpy(&x, &v, 4); floats.push_back(x); }
                break;
            case 7: if (f.wire == 2) packed_ints(f.bytes, ints); else ints.push_back(static_cast<int64_t>(f.value)); break;
            case 8: t.name = f.bytes.str(); break;
            case 9: raw = f.bytes; has_raw = true; break;
            default: break;
        }
    }
    // TensorProto.DataType: 1 float, 6 int32, 7 int64, 9 bool, 10 float16, 11 double
    if (has_raw) {
        const uint8_t* p = raw.p;
        switch (t.dtype) {
            case 1: t.f.resize(raw.n / 4); std::memcpy(t.f.data(), p, t.f.size() * 4); break;
            case 10: t.f.resize(raw.n / 2); for (size_t k = 0; k < t.f.size(); ++k) { uint16_t h; std::memcpy(&h, p + 2 * k, 2); t.f[k] = half_bits_to_float(h); } break;
            case 11: t.f.resize(raw.n / 8); for (size_t k = 0; k < t.f.size(); ++k) { double d; std::memcpy(&d, p + 8 * k, 8); t.f[k] = static_cast<float>(d); } break;
            case 7: t.i.resize(raw.n / 8); std::memcpy(t.i.data(), p, t.i.size() * 8); break;
            case 6: t.i.resize(raw.n / 4); for (size_t k = 0; k < t.i.size(); ++k) { int32_t v; std::memcpy(&v, p + 4 * k, 4); t.i[k] = v; } break;
            case 9: t.i.resize(raw.n); for (size_t k = 0; k < raw.n; ++k) t.i[k] = p[k]; break;
            default: break;   // types the plan never reads
        }
    } else if (t.dtype == 1 || t.dtype == 10 || t.dtype == 11) {
        t.f = std::move(floats);
    } else {
        t.i = std::move(ints);
    }
    return t;
}

struct Node {
    std::string op, name;
    std::vector<std::string> in, out;
    std::map<std::string, std::vector<int64_t>> ints;   // INT and INTS attributes
    std::map<std::string, std::string> strs;            // STRING attributes
    int64_t int_attr(const char* k, int64_t dflt) const {
        const auto it = ints.find(k);
        return it == ints.end() || it->second.empty() ? dflt : it->second[0];
    }
    const std::vector<int64_t>& ints_attr(const char* k) const {
        const auto it = ints.find(k);
        if (it == ints.end()) throw std::runtime_error("onnx: node " + name + " lacks attribute " + k);
        return it->second;
    }
};

Node parse_node(Span s) {
    Node n;
    for (Fields f(s); f.next();) {
        switch (f.field) {
            case 1: n.in.push_back(f.bytes.str()); break;
            case 2: n.out.push_back(f.bytes.str()); break;
            case 3: n.name = f.bytes.str(); break;
            case 4: n.op = f.bytes.str(); break;
            case 5: {   // AttributeProto: 1 name, 3 i, 4 s, 8 ints
                std::string an, sv;
                std::vector<int64_t> iv;
                bool has_i = false, has_s = false;
                int64_t i = 0;
                for (Fields a(f.bytes); a.next();) {
                    if (a.field == 1) an = a.bytes.str();
                    else if (a.field == 3) { i = static_cast<int64_t>(a.value); has_i = true; }
                    else if (a.field == 4) { sv = a.bytes.str(); has_s = true; }
                    else if (a.field == 8) { if (a.wire == 2) packed_ints(a.bytes, iv); else iv.push_back(static_cast<int64_t>(a.value)); }
                }
                if (!iv.empty()) n.ints[an] = iv;
                else if (has_i) n.ints[an] = {i};
                if (has_s) n.strs[an] = sv;
                break;
            }
            default: break;
        }
    }
    return n;
}

struct Graph {
    std::vector<Node> nodes;
    std::unordered_map<std::string, Tensor> init;
    std::string input;   // first graph input that is not an initializer
};

Graph parse_model(const std::vector<uint8_t>& file, const std::string& path) {
    Span graph;
    for (Fields f(Span{file.data(), file.size()}); f.next();)
        if (f.field == 7 && f.wire == 2) graph = f.bytes;   // ModelProto.graph
    if (!graph.p) throw std::runtime_error(path + ": no GraphProto");
    Graph g;
    std::vector<std::string> inputs;
    for (Fields f(graph); f.next();) {
        if (f.wire != 2) continue;
        if (f.field == 1) g.nodes.push_back(parse_node(f.bytes));
        else if (f.field == 5) { Tensor t = parse_tensor(f.bytes); std::string k = t.name; g.init.emplace(std::move(k), std::move(t)); }
        else if (f.field == 11)
            for (Fields v(f.bytes); v.next();)
                if (v.field == 1) inputs.push_back(v.bytes.str());
    }
    for (const std::string& n : inputs)
        if (!g.init.count(n)) { g.input = n; break; }
    if (g.input.empty()) throw std::runtime_error(path + ": graph has no input");
    return g;
}

// ------------------------------------------------------------------------------------------------------------------
// graph -> plan
// ------------------------------------------------------------------------------------------------------------------
enum { OP_CONV = 0, OP_MAXPOOL5 = 1, OP_UPSAMPLE2 = 2, OP_COPY = 3 };
enum { DT_F16 = 0, DT_F32 = 1 };

struct Shape { int c, h, w; };
struct View { int buf, coff, c, h, w; };
struct PendingCopy { const Node* concat; int input; int buf, off; };

int align_up(int x, int a) { return (x + a - 1) / a * a; }

class Compiler {
public:
    Compiler(const Graph& g, int in_h, int in_w) : g_(g), in_h_(in_h), in_w_(in_w) {
        for (const Node& n : g.nodes) {
            for (const std::string& i : n.in) consumers_[i].push_back(&n);
            for (const std::string& o : n.out) producer_[o] = &n;
        }
    }

    std::vector<uint8_t> run() {
        find_levels();
        feature_part();
        infer_shapes();
        home_concats();
        emit();
        return serialize();
    }

private:
    const Node* producer(const std::string& t) const {
        const auto it = producer_.find(t);
        return it == producer_.end() ? nullptr : it->second;
    }
    const std::vector<const Node*>& consumers(const std::string& t) const {
        static const std::vector<const Node*> none;
        const auto it = consumers_.find(t);
        return it == consumers_.end() ? none : it->second;
    }
    const Tensor& init(const std::string& t) const {
        const auto it = g_.init.find(t);
        if (it == g_.init.end()) throw std::runtime_error("onnx: initializer expected: " + t);
        return it->second;
    }
    const Shape& shape(const std::string& t) const {
        const auto it = shapes_.find(t);
        if (it == shapes_.end()) throw std::runtime_error("onnx: no shape for " + t);
        return it->second;
    }
    const View& view(const std::string& t) const {
        const auto it = views_.find(t);
        if (it == views_.end()) throw std::runtime_error("onnx: tensor used before it is produced: " + t);
        return it->second;
    }
    int new_buf(int h, int w, int c, int dtype) {
        bufs_.push_back(EngineBuf{h, w, c, dtype});
        return static_cast<int>(bufs_.size()) - 1;
    }

    // the per-level Concat([box conv, class conv]) that feeds a Reshape; levels in the anchor order of the axis-2 Concat
    void find_levels() {
        std::vector<const Node*> found;
        for (const Node& n : g_.nodes) {
            if (n.op != "Concat" || n.int_attr("axis", 0) != 1) continue;
            bool to_reshape = false, from_convs = true;
            for (const Node* c : consumers(n.out[0])) to_reshape |= c->op == "Reshape";
            for (const std::string& i : n.in) { const Node* p = producer(i); from_convs &= p && p->op == "Conv"; }
            if (to_reshape && from_convs) found.push_back(&n);
        }
        if (found.empty()) throw std::runtime_error("onnx: no detection head found");
        const Node *reshape = nullptr, *cat2 = nullptr;
        for (const Node* c : consumers(found[0]->out[0])) if (c->op == "Reshape") { reshape = c; break; }
        for (const Node* c : consumers(reshape->out[0])) if (c->op == "Concat") { cat2 = c; break; }
        if (!cat2) throw std::runtime_error("onnx: head levels are not concatenated");
        for (const std::string& r : cat2->in) {
            const Node* p = producer(r);
            const Node* level = nullptr;
            for (const Node* n : found) if (p && n->out[0] == p->in[0]) { level = n; break; }
            if (!level) throw std::runtime_error("onnx: unexpected head layout");
            levels_.push_back(level);
        }
    }

    // nodes up to (and including) the one that produces the last level Concat
    void feature_part() {
        std::set<std::string> want, have;
        for (const Node* n : levels_) want.insert(n->out[0]);
        for (const Node& n : g_.nodes) {
            feature_.push_back(&n);
            for (const std::string& o : n.out) if (want.count(o)) have.insert(o);
            if (have.size() == want.size()) break;
        }
    }

    void infer_shapes() {
        shapes_[g_.input] = Shape{3, in_h_, in_w_};
        for (const Node* n : feature_) {
            if (n->op == "Conv") {
                const Shape s = shape(n->in[0]);
                const Tensor& w = init(n->in[1]);
                const int st = static_cast<int>(n->ints_attr("strides")[0]), k = static_cast<int>(n->ints_attr("kernel_shape")[0]),
                          p = static_cast<int>(n->ints_attr("pads")[0]);
                shapes_[n->out[0]] = Shape{static_cast<int>(w.dims.at(0)), (s.h + 2 * p - k) / st + 1, (s.w + 2 * p - k) / st + 1};
            } else if (n->op == "Sigmoid" || n->op == "Mul" || n->op == "Add" || n->op == "MaxPool") {
                const Shape* s = nullptr;
                for (const std::string& i : n->in) if (shapes_.count(i)) { s = &shapes_[i]; break; }
                if (!s) throw std::runtime_error("onnx: no shaped input for " + n->name);
                shapes_[n->out[0]] = *s;
            } else if (n->op == "Split") {
                const Shape s = shape(n->in[0]);
                const Tensor& sizes = init(n->in[1]);
                for (size_t o = 0; o < n->out.size() && o < sizes.i.size(); ++o)
                    shapes_[n->out[o]] = Shape{static_cast<int>(sizes.i[o]), s.h, s.w};
            } else if (n->op == "Concat") {
                if (n->int_attr("axis", 0) != 1) throw std::runtime_error("onnx: Concat over a non-channel axis in the feature part");
                int c = 0;
                for (const std::string& i : n->in) c += shape(i).c;
                shapes_[n->out[0]] = Shape{c, shape(n->in[0]).h, shape(n->in[0]).w};
            } else if (n->op == "Resize") {
                const Shape s = shape(n->in[0]);
                shapes_[n->out[0]] = Shape{s.c, 2 * s.h, 2 * s.w};
            } else if (n->op == "Slice") {
                const Shape s = shape(n->in[0]);
                const int64_t st = init(n->in[1]).first_int(), en = init(n->in[2]).first_int(), ax = init(n->in[3]).first_int();
                if (ax != 1 || (n->in.size() >= 5 && init(n->in[4]).first_int() != 1))
                    throw std::runtime_error("onnx: only unit-step channel Slice is supported");
                shapes_[n->out[0]] = Shape{static_cast<int>(std::min<int64_t>(en, s.c) - st), s.h, s.w};
            } else {
                throw std::runtime_error("onnx: " + n->op + " in feature part (" + n->name + ")");
            }
        }
    }

    // Concat costs nothing when each producer writes straight into its channel range of the concat buffer ("home").
    // A Split whose outputs all reappear, in order, in one Concat homes the Split's input instead; what cannot be
    // homed (a tensor that already lives elsewhere, a Slice) is copied.
    void home_concats() {
        for (const Node* n : feature_) {
            if (n->op != "Concat") continue;
            const Shape s = shape(n->out[0]);
            const bool is_level = std::find(levels_.begin(), levels_.end(), n) != levels_.end();
            const int b = new_buf(s.h, s.w, is_level ? align_up(s.c, 4) : align_up(s.c, 8), is_level ? DT_F32 : DT_F16);
            views_[n->out[0]] = View{b, 0, s.c, s.h, s.w};
            int off = 0;
            for (size_t i = 0; i < n->in.size();) {
                const std::string& name = n->in[i];
                const Node* pr = producer(name);
                if (pr && pr->op == "Split") {
                    const std::vector<std::string>& outs = pr->out;
                    const bool run = i + outs.size() <= n->in.size() && std::equal(outs.begin(), outs.end(), n->in.begin() + i);
                    if (run && !home_.count(pr->in[0]) && consumers(pr->in[0]).size() == 1) {
                        home_[pr->in[0]] = {b, off};
                        for (const std::string& o : outs) off += shape(o).c;
                        i += outs.size();
                        continue;
                    }
                    pending_.push_back(PendingCopy{n, static_cast<int>(i), b, off});
                } else if (pr && pr->op == "Slice") {
                    pending_.push_back(PendingCopy{n, static_cast<int>(i), b, off});
                } else if (home_.count(name) || views_.count(name)) {
                    pending_.push_back(PendingCopy{n, static_cast<int>(i), b, off});
                } else {
                    home_[name] = {b, off};
                }
                off += shape(name).c;
                ++i;
            }
        }
    }

    View place(const std::string& name) {
        const Shape s = shape(name);
        const auto it = home_.find(name);
        const View v = it != home_.end() ? View{it->second.first, it->second.second, s.c, s.h, s.w}
                                         : View{new_buf(s.h, s.w, align_up(s.c, 8), DT_F16), 0, s.c, s.h, s.w};
        views_[name] = v;
        return v;
    }

    // weights [Cout_pad][tap][Cin_pad] fp16 (K-major B operand), bias [Cout_pad] fp32
    void add_weights(const Tensor& w, const Tensor* bias, int cin_pad, EngineOp& op) {
        const int cout = static_cast<int>(w.dims.at(0)), cin = static_cast<int>(w.dims.at(1)),
                  taps = static_cast<int>(w.dims.at(2) * w.dims.at(3));
        const int cout_pad = align_up(cout, 16);
        if (w.f.size() != static_cast<size_t>(cout) * cin * taps) throw std::runtime_error("onnx: weight tensor size mismatch: " + w.name);
        blob_.resize(align_up64(blob_.size(), 1024), 0);
        op.w_off = static_cast<int64_t>(blob_.size());
        blob_.resize(blob_.size() + static_cast<size_t>(cout_pad) * taps * cin_pad * 2, 0);
        uint16_t* dst = reinterpret_cast<uint16_t*>(blob_.data() + op.w_off);
        for (int o = 0; o < cout; ++o)
            for (int c = 0; c < cin; ++c)
                for (int t = 0; t < taps; ++t)
                    dst[(static_cast<size_t>(o) * taps + t) * cin_pad + c] = float_to_half_bits(w.f[(static_cast<size_t>(o) * cin + c) * taps + t]);
        blob_.resize(align_up64(blob_.size(), 256), 0);
        op.b_off = static_cast<int64_t>(blob_.size());
        blob_.resize(blob_.size() + static_cast<size_t>(cout_pad) * 4, 0);
        if (bias) std::memcpy(blob_.data() + op.b_off, bias->f.data(), std::min<size_t>(bias->f.size(), cout) * 4);
        op.cout_pad = cout_pad;
        op.cin_pad = cin_pad;
    }
    static size_t align_up64(size_t x, size_t a) { return (x + a - 1) / a * a; }

    static EngineOp make_op(int type, const View& s, const View& d) {
        EngineOp op{};
        op.type = type;
        op.src_buf = s.buf; op.src_coff = s.coff; op.src_c = s.c; op.src_h = s.h; op.src_w = s.w;
        op.dst_buf = d.buf; op.dst_coff = d.coff; op.dst_c = d.c; op.dst_h = d.h; op.dst_w = d.w;
        op.k = 1; op.stride = 1; op.res_buf = -1;
        return op;
    }

    void emit() {
        views_[g_.input] = View{new_buf(in_h_, in_w_, 4, DT_F16), 0, 3, in_h_, in_w_};   // NHWC fp16, 3 channels padded to 4
        input_buf_ = views_[g_.input].buf;
        std::set<const Node*> fused;
        for (const Node* n : feature_) {
            if (fused.count(n)) continue;
            if (n->op == "Conv") {
                // Conv -> (Sigmoid, Mul) = SiLU -> (Add with a non-constant operand) = bottleneck shortcut
                std::string out = n->out[0], res;
                int act = 0;
                const auto& c1 = consumers(out);
                if (c1.size() == 2 && ((c1[0]->op == "Sigmoid" && c1[1]->op == "Mul") || (c1[0]->op == "Mul" && c1[1]->op == "Sigmoid"))) {
                    const Node* mul = c1[0]->op == "Mul" ? c1[0] : c1[1];
                    const Node* sig = c1[0]->op == "Mul" ? c1[1] : c1[0];
                    const bool silu = mul->in.size() == 2 && ((mul->in[0] == out && mul->in[1] == sig->out[0]) || (mul->in[1] == out && mul->in[0] == sig->out[0]));
                    if (!silu) throw std::runtime_error("onnx: Sigmoid/Mul pair is not x*sigmoid(x) at " + n->name);
                    fused.insert(mul); fused.insert(sig);
                    out = mul->out[0];
                    act = 1;
                    const auto& c2 = consumers(out);
                    if (c2.size() == 1 && c2[0]->op == "Add") {
                        std::vector<std::string> other;
                        for (const std::string& i : c2[0]->in) if (i != out) other.push_back(i);
                        if (other.size() == 1 && !g_.init.count(other[0])) {
                            res = other[0];
                            fused.insert(c2[0]);
                            out = c2[0]->out[0];
                        }
                    }
                }
                shapes_[out] = shape(n->out[0]);
                const View src = view(n->in[0]);
                const View dst = place(out);
                const Tensor& w = init(n->in[1]);
                const int cin = static_cast<int>(w.dims.at(1)), k = static_cast<int>(n->ints_attr("kernel_shape")[0]);
                if (cin != src.c) throw std::runtime_error("onnx: channel mismatch at " + n->name);
                if (n->ints_attr("pads")[0] != k / 2 || n->int_attr("group", 1) != 1)
                    throw std::runtime_error("onnx: only 'same' padded, ungrouped convolutions are supported (" + n->name + ")");
                EngineOp op = make_op(OP_CONV, src, dst);
                op.k = k;
                op.stride = static_cast<int>(n->ints_attr("strides")[0]);
                op.act = act;
                if (!res.empty()) { op.res_buf = view(res).buf; op.res_coff = view(res).coff; }
                add_weights(w, n->in.size() > 2 ? &init(n->in[2]) : nullptr, cin == 3 ? 4 : cin, op);
                ops_.push_back(op);
            } else if (n->op == "Split") {
                const View v = view(n->in[0]);
                int off = 0;
                for (const std::string& o : n->out) {
                    views_[o] = View{v.buf, v.coff + off, shape(o).c, v.h, v.w};
                    off += shape(o).c;
                }
            } else if (n->op == "Slice") {
                const View v = view(n->in[0]);
                views_[n->out[0]] = View{v.buf, v.coff + static_cast<int>(init(n->in[1]).first_int()), shape(n->out[0]).c, v.h, v.w};
            } else if (n->op == "MaxPool") {
                const auto &ks = n->ints_attr("kernel_shape"), &st = n->ints_attr("strides");
                if (ks != std::vector<int64_t>{5, 5} || st != std::vector<int64_t>{1, 1} || n->ints_attr("pads")[0] != 2)
                    throw std::runtime_error("onnx: only the SPPF 5x5 stride-1 MaxPool is supported");
                const View src = view(n->in[0]);
                ops_.push_back(make_op(OP_MAXPOOL5, src, place(n->out[0])));
            } else if (n->op == "Resize") {
                const auto it = n->strs.find("mode");
                if (it == n->strs.end() || it->second != "nearest") throw std::runtime_error("onnx: only nearest Resize is supported");
                const View src = view(n->in[0]);
                ops_.push_back(make_op(OP_UPSAMPLE2, src, place(n->out[0])));
            } else if (n->op == "Concat") {
                for (const PendingCopy& pc : pending_) {
                    if (pc.concat != n) continue;
                    const View sv = view(n->in[pc.input]);
                    ops_.push_back(make_op(OP_COPY, sv, View{pc.buf, pc.off, sv.c, sv.h, sv.w}));
                }
            } else if (n->op == "Sigmoid" || n->op == "Mul" || n->op == "Add") {
                throw std::runtime_error("onnx: unfused " + n->op + " " + n->name);
            }
        }
        for (const Node* n : levels_)
            if (shape(n->in[0]).c != 64) throw std::runtime_error("onnx: DFL head with reg_max = 16 expected");
        num_classes_ = shape(levels_[0]->in[1]).c;
    }

    std::vector<uint8_t> serialize() const {
        struct Header {
            char magic[8];
            int32_t n_bufs, n_ops, n_levels, num_classes, in_h, in_w, input_buf, reserved;
            int64_t blob_bytes;
        } h{};
        static_assert(sizeof(Header) == 48, "engine header layout");
        std::memcpy(h.magic, "RMRENG2", 8);
        h.n_bufs = static_cast<int32_t>(bufs_.size());
        h.n_ops = static_cast<int32_t>(ops_.size());
        h.n_levels = static_cast<int32_t>(levels_.size());
        h.num_classes = num_classes_;
        h.in_h = in_h_; h.in_w = in_w_; h.input_buf = input_buf_;
        h.blob_bytes = static_cast<int64_t>(blob_.size());
        std::vector<uint8_t> out;
        auto put = [&out](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); out.insert(out.end(), b, b + n); };
        put(&h, sizeof(h));
        put(bufs_.data(), sizeof(EngineBuf) * bufs_.size());
        put(ops_.data(), sizeof(EngineOp) * ops_.size());
        for (const Node* n : levels_) {
            const View& v = view(n->out[0]);
            const int32_t rec[4] = {v.buf, v.h, v.w, in_h_ / v.h};
            put(rec, sizeof(rec));
        }
        out.resize(align_up64(out.size(), 1024), 0);
        put(blob_.data(), blob_.size());
        return out;
    }

    const Graph& g_;
    int in_h_, in_w_, input_buf_ = 0, num_classes_ = 0;
    std::unordered_map<std::string, std::vector<const Node*>> consumers_;
    std::unordered_map<std::string, const Node*> producer_;
    std::vector<const Node*> levels_, feature_;
    std::unordered_map<std::string, Shape> shapes_;
    std::unordered_map<std::string, View> views_;
    std::unordered_map<std::string, std::pair<int, int>> home_;
    std::vector<PendingCopy> pending_;
    std::vector<EngineBuf> bufs_;
    std::vector<EngineOp> ops_;
    std::vector<uint8_t> blob_;
};

bool exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

std::vector<uint8_t> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::invalid_argument("cannot open " + path);
    f.seekg(0, std::ios::end);
    std::vector<uint8_t> data(static_cast<size_t>(f.tellg()));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(data.data()), static_cast<std::streamsize>(data.size()));
    if (!f) throw std::runtime_error("short read: " + path);
    return data;
}

}  // namespace

std::vector<uint8_t> compile_onnx(const std::string& onnx_path, int in_h, int in_w) {
    const std::vector<uint8_t> file = read_file(onnx_path);
    const Graph g = parse_model(file, onnx_path);
    return Compiler(g, in_h, in_w).run();
}

void build_engine(const std::string& onnx_path, const std::string& engine_path, int in_h, int in_w) {
    const std::vector<uint8_t> bytes = compile_onnx(onnx_path, in_h, in_w);
    // write beside, then rename: a reader never sees a half-written plan (the reference writes in place,
    // detector.cpp:281-311, and relies on one process building)
    const std::string tmp = engine_path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        if (!f) throw std::runtime_error("cannot write " + tmp);
        f.write(reinterpret_cast<const char*>(bytes.data()), static_cast<std::streamsize>(bytes.size()));
        if (!f) throw std::runtime_error("short write: " + tmp);
    }
    if (std::rename(tmp.c_str(), engine_path.c_str()) != 0) {
        std::remove(tmp.c_str());
        throw std::runtime_error("cannot move " + tmp + " to " + engine_path);
    }
}

std::string resolve_engine(const std::string& path, int in_h, int in_w) {
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash) && dot + 1 < path.size();
    const std::string base = has_ext ? path.substr(0, dot) : path;
    const std::string ext = has_ext ? path.substr(dot) : "";
    const std::string eng = ext == ".rmeng" ? path : base + ".rmeng";
    if (exists(eng)) return eng;
    const std::string onnx = base + ".onnx";
    // detector.cpp:80: neither the engine nor the ONNX file -> std::invalid_argument
    if (!exists(onnx)) throw std::invalid_argument("neither " + eng + " nor " + onnx + " exists");
    try {
        build_engine(onnx, eng, in_h, in_w);
        return eng;
    } catch (const std::runtime_error&) {
        // read-only model directory: cache under $RMR_ENGINE_CACHE when the caller named one
        const char* cache = std::getenv("RMR_ENGINE_CACHE");
        if (!cache || !cache[0]) throw;
        const std::string alt = std::string(cache) + "/" + (slash == std::string::npos ? base : base.substr(slash + 1)) + ".rmeng";
        if (!exists(alt)) build_engine(onnx, alt, in_h, in_w);
        return alt;
    }
}

}  // namespace rmr
