"""Engine builder: ONNX graph → static layer plan + packed fp16 weights (`.rmeng` file).

Replaces the reference's TensorRT engine build/cache step
(`/root/reference/src/detect/detector.cpp:70-99,177-243,281-311`): when `<name>.engine` is absent the
reference parses the sibling `<name>.onnx`, builds an FP16 engine and writes it next to it.  Here
the "engine" is a flat binary the C++ runtime (`csrc/net.cu`) maps 1:1 onto sm_100a kernel launches:

  header | buffers[] | ops[] | levels[] | weight blob (fp16 weights, fp32 bias)

Compile-time fusions (done here, once, offline):
  * Conv + bias + Sigmoid·Mul (SiLU) → one CONV op with act=1
  * bottleneck shortcut `Add(x, conv(...))` → residual operand of the producing CONV
  * Concat → no op: every producer writes at its channel offset inside the concat buffer
  * Split  → no op: consumers read a channel-offset view
  * head tail (reshape / DFL softmax / dist2bbox / sigmoid / concat) → one DECODE stage in
    `csrc/postprocess.cu`; the plan only records the per-level [H,W,64+nc] fp32 logits buffers

Activations are NHWC fp16 with a per-buffer channel pitch; weights are [Cout_pad][tap][Cin] fp16
(K-major, the layout the tcgen05 B operand wants), bias fp32.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from . import onnx_wire

MAGIC = b"RMRENG2\0"

OP_CONV = 0       # tcgen05 implicit GEMM (or SIMT stem when cin_pad == 4)
OP_MAXPOOL5 = 1
OP_UPSAMPLE2 = 2
OP_COPY = 3

DT_F16 = 0
DT_F32 = 1


@dataclass
class Buf:
    H: int
    W: int
    C: int          # channel pitch
    dtype: int = DT_F16


@dataclass
class View:
    buf: int
    coff: int
    C: int
    H: int
    W: int


@dataclass
class Op:
    type: int
    src: View
    dst: View
    k: int = 1
    stride: int = 1
    act: int = 0
    res: View | None = None
    w_off: int = 0
    b_off: int = 0
    cout_pad: int = 0
    cin_pad: int = 0
    name: str = ""


@dataclass
class Plan:
    bufs: list = field(default_factory=list)
    ops: list = field(default_factory=list)
    levels: list = field(default_factory=list)   # (buf, H, W, stride)
    num_classes: int = 0
    in_h: int = 640
    in_w: int = 640
    blob: bytearray = field(default_factory=bytearray)
    macs: int = 0


def _align(x, a):
    return (x + a - 1) // a * a


def compile_onnx(path: str, in_h: int = 640, in_w: int = 640) -> Plan:
    g = onnx_wire.load(path)
    init = g.initializers
    consumers: dict[str, list] = {}
    producer: dict[str, onnx_wire.Node] = {}
    for n in g.nodes:
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)
        for o in n.outputs:
            producer[o] = n

    plan = Plan(in_h=in_h, in_w=in_w)
    views: dict[str, View] = {}
    home: dict[str, tuple] = {}       # tensor name -> (buf, coff)

    # ---- which nodes belong to the feature part (everything before the per-level Reshape) ----
    level_concats = []
    for n in g.nodes:
        if n.op == "Concat" and n.attrs.get("axis") == 1:
            cons = consumers.get(n.outputs[0], [])
            if any(c.op == "Reshape" for c in cons) and all(producer[i].op == "Conv" for i in n.inputs):
                level_concats.append(n)
    assert len(level_concats) >= 1, "no detection head found"
    # anchor order = order of the axis-2 Concat over the reshaped levels
    first_reshape = [c for c in consumers[level_concats[0].outputs[0]] if c.op == "Reshape"][0]
    cat2 = [c for c in consumers[first_reshape.outputs[0]] if c.op == "Concat"][0]
    order = []
    for rname in cat2.inputs:
        src = producer[rname].inputs[0]
        order.append([n for n in level_concats if n.outputs[0] == src][0])
    level_concats = order

    # ---- shape inference for the feature part (only needs H, W, C) ----
    shapes: dict[str, tuple] = {g.inputs[0].name: (3, in_h, in_w)}

    def final_name(conv_node):
        """Follow Conv → (Sigmoid, Mul) → (Add) fusion; returns (result tensor, act, residual name)."""
        out = conv_node.outputs[0]
        act = 0
        res = None
        cons = consumers.get(out, [])
        if len(cons) == 2 and {c.op for c in cons} == {"Sigmoid", "Mul"}:
            mul = [c for c in cons if c.op == "Mul"][0]
            sig = [c for c in cons if c.op == "Sigmoid"][0]
            assert set(mul.inputs) == {out, sig.outputs[0]}
            out = mul.outputs[0]
            act = 1
            c2 = consumers.get(out, [])
            if len(c2) == 1 and c2[0].op == "Add":
                add = c2[0]
                other = [i for i in add.inputs if i != out]
                if len(other) == 1 and other[0] not in init:
                    res = other[0]
                    out = add.outputs[0]
        return out, act, res

    stop = {n.outputs[0] for n in level_concats}
    feature_nodes = []
    for n in g.nodes:
        feature_nodes.append(n)
        if stop.issubset(set().union(*[set(m.outputs) for m in feature_nodes])):
            break

    # ---- pass 1: concat homes ----
    def new_buf(H, W, C, dtype=DT_F16):
        plan.bufs.append(Buf(H, W, C, dtype))
        return len(plan.bufs) - 1

    # shapes first (cheap symbolic run)
    for n in feature_nodes:
        if n.op == "Conv":
            c, h, w = shapes[n.inputs[0]]
            wt = init[n.inputs[1]]
            s = n.attrs["strides"][0]
            k = n.attrs["kernel_shape"][0]
            p = n.attrs["pads"][0]
            shapes[n.outputs[0]] = (wt.shape[0], (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1)
        elif n.op in ("Sigmoid", "Mul", "Add", "MaxPool"):
            shapes[n.outputs[0]] = shapes[[i for i in n.inputs if i in shapes][0]]
        elif n.op == "Split":
            c, h, w = shapes[n.inputs[0]]
            for o, sz in zip(n.outputs, init[n.inputs[1]].tolist()):
                shapes[o] = (int(sz), h, w)
        elif n.op == "Concat":
            assert n.attrs["axis"] == 1
            ss = [shapes[i] for i in n.inputs]
            shapes[n.outputs[0]] = (sum(s[0] for s in ss), ss[0][1], ss[0][2])
        elif n.op == "Resize":
            c, h, w = shapes[n.inputs[0]]
            shapes[n.outputs[0]] = (c, 2 * h, 2 * w)
        elif n.op == "Slice":
            c, h, w = shapes[n.inputs[0]]
            st, en, ax = (int(init[n.inputs[j]].reshape(-1)[0]) for j in (1, 2, 3))
            assert ax == 1 and (len(n.inputs) < 5 or int(init[n.inputs[4]].reshape(-1)[0]) == 1)
            en = min(en, c)
            shapes[n.outputs[0]] = (en - st, h, w)
        else:
            raise NotImplementedError(f"{n.op} in feature part ({n.name})")

    pending_copies = []   # (concat node, input index) that could not be homed
    for n in feature_nodes:
        if n.op != "Concat":
            continue
        C, H, W = shapes[n.outputs[0]]
        is_level = n in level_concats
        pitch = _align(C, 4) if is_level else _align(C, 8)
        b = new_buf(H, W, pitch, DT_F32 if is_level else DT_F16)
        views[n.outputs[0]] = View(b, 0, C, H, W)
        off = 0
        i = 0
        while i < len(n.inputs):
            name = n.inputs[i]
            c = shapes[name][0]
            pr = producer.get(name)
            if pr is not None and pr.op == "Split":
                outs = pr.outputs
                if n.inputs[i:i + len(outs)] == outs and pr.inputs[0] not in home and \
                        len(consumers[pr.inputs[0]]) == 1:
                    home[pr.inputs[0]] = (b, off)
                    tot = sum(shapes[o][0] for o in outs)
                    off += tot
                    i += len(outs)
                    continue
                pending_copies.append((n, i, b, off))
            elif pr is not None and pr.op == "Slice":
                pending_copies.append((n, i, b, off))
            elif name in home or name in views:
                pending_copies.append((n, i, b, off))
            else:
                home[name] = (b, off)
            off += c
            i += 1

    # ---- pass 2: emit ops ----
    def place(name):
        C, H, W = shapes[name]
        if name in home:
            b, off = home[name]
            v = View(b, off, C, H, W)
        else:
            v = View(new_buf(H, W, _align(C, 8)), 0, C, H, W)
        views[name] = v
        return v

    # network input: NHWC fp16, 3 channels padded to 4
    in_name = g.inputs[0].name
    views[in_name] = View(new_buf(in_h, in_w, 4), 0, 3, in_h, in_w)
    plan.input_buf = views[in_name].buf

    def add_weights(wt: np.ndarray, bias: np.ndarray | None, cin_pad: int):
        cout, cin, kh, kw = wt.shape
        cout_pad = _align(cout, 16)
        w = np.zeros((cout_pad, kh * kw, cin_pad), np.float16)
        w[:cout, :, :cin] = wt.transpose(0, 2, 3, 1).reshape(cout, kh * kw, cin).astype(np.float16)
        b = np.zeros(cout_pad, np.float32)
        if bias is not None:
            b[:cout] = bias
        while len(plan.blob) % 1024:
            plan.blob.append(0)
        w_off = len(plan.blob)
        plan.blob += w.tobytes()
        while len(plan.blob) % 256:
            plan.blob.append(0)
        b_off = len(plan.blob)
        plan.blob += b.tobytes()
        return w_off, b_off, cout_pad

    skip = set()
    for n in feature_nodes:
        if n.name in skip:
            continue
        if n.op == "Conv":
            out, act, res = final_name(n)
            # mark fused nodes
            t = n.outputs[0]
            while t != out:
                for c in consumers[t]:
                    skip.add(c.name)
                nxt = [c for c in consumers[t] if c.op in ("Mul", "Add")]
                t = nxt[0].outputs[0]
            shapes[out] = shapes[n.outputs[0]]
            src = views[n.inputs[0]]
            dst = place(out)
            wt = init[n.inputs[1]]
            bias = init[n.inputs[2]] if len(n.inputs) > 2 else None
            cin = wt.shape[1]
            assert cin == src.C
            cin_pad = 4 if cin == 3 else cin
            w_off, b_off, cout_pad = add_weights(wt, bias, cin_pad)
            k = n.attrs["kernel_shape"][0]
            s = n.attrs["strides"][0]
            assert n.attrs["pads"][0] == k // 2 and n.attrs["group"] == 1
            plan.ops.append(Op(OP_CONV, src, dst, k, s, act, views[res] if res else None,
                               w_off, b_off, cout_pad, cin_pad, n.name))
            plan.macs += wt.size * dst.H * dst.W
        elif n.op == "Split":
            v = views[n.inputs[0]]
            off = 0
            for o in n.outputs:
                c = shapes[o][0]
                views[o] = View(v.buf, v.coff + off, c, v.H, v.W)
                off += c
        elif n.op == "Slice":
            v = views[n.inputs[0]]
            st = int(init[n.inputs[1]].reshape(-1)[0])
            views[n.outputs[0]] = View(v.buf, v.coff + st, shapes[n.outputs[0]][0], v.H, v.W)
        elif n.op == "MaxPool":
            assert n.attrs["kernel_shape"] == [5, 5] and n.attrs["strides"] == [1, 1] and n.attrs["pads"][0] == 2
            plan.ops.append(Op(OP_MAXPOOL5, views[n.inputs[0]], place(n.outputs[0]), name=n.name))
        elif n.op == "Resize":
            assert n.attrs["mode"] == "nearest"
            plan.ops.append(Op(OP_UPSAMPLE2, views[n.inputs[0]], place(n.outputs[0]), name=n.name))
        elif n.op == "Concat":
            for (cn, i, b, off) in pending_copies:
                if cn is n:
                    sv = views[n.inputs[i]]
                    plan.ops.append(Op(OP_COPY, sv, View(b, off, sv.C, sv.H, sv.W), name=n.name))
        elif n.op in ("Sigmoid", "Mul", "Add"):
            raise NotImplementedError(f"unfused {n.op} {n.name}")

    for n in level_concats:
        v = views[n.outputs[0]]
        plan.levels.append((v.buf, v.H, v.W, in_h // v.H))
        box_c = shapes[n.inputs[0]][0]
        assert box_c == 64, "DFL head with reg_max=16 expected"
    plan.num_classes = shapes[level_concats[0].inputs[1]][0]
    return plan


_OP_FMT = "<" + "i" * 24 + "qq"


def serialize(plan: Plan) -> bytes:
    out = bytearray()
    out += MAGIC
    out += struct.pack("<iiiiiiiiq", len(plan.bufs), len(plan.ops), len(plan.levels), plan.num_classes,
                       plan.in_h, plan.in_w, plan.input_buf, 0, len(plan.blob))
    for b in plan.bufs:
        out += struct.pack("<iiii", b.H, b.W, b.C, b.dtype)
    for op in plan.ops:
        r = op.res
        out += struct.pack(
            _OP_FMT, op.type,
            op.src.buf, op.src.coff, op.src.C, op.src.H, op.src.W,
            op.dst.buf, op.dst.coff, op.dst.C, op.dst.H, op.dst.W,
            op.k, op.stride, op.act,
            r.buf if r else -1, r.coff if r else 0,
            op.cout_pad, op.cin_pad, 0, 0, 0, 0, 0, 0,
            op.w_off, op.b_off)
    for (b, H, W, s) in plan.levels:
        out += struct.pack("<iiii", b, H, W, s)
    while len(out) % 1024:
        out.append(0)
    out += plan.blob
    return bytes(out)


def build_engine(onnx_path: str, engine_path: str, in_h: int = 640, in_w: int = 640) -> Plan:
    plan = compile_onnx(onnx_path, in_h, in_w)
    with open(engine_path, "wb") as fh:
        fh.write(serialize(plan))
    return plan


if __name__ == "__main__":
    import sys
    p = build_engine(sys.argv[1], sys.argv[2])
    print(f"{sys.argv[2]}: {len(p.ops)} ops, {len(p.bufs)} buffers, {len(p.blob) / 1e6:.1f} MB weights, "
          f"{p.macs / 1e9:.3f} GMAC, classes={p.num_classes}, levels={p.levels}")
