#!/usr/bin/env python
"""TensorRT stand-in: the reference executes car.onnx / armor.onnx through TensorRT with the FP16 flag
(/root/reference/src/detect/detector.h:122, detector.cpp:223-231).  TensorRT is not in this image, so the
library path on the same B200 is measured with what is: the same two ONNX graphs run op by op through
torch + cuDNN in fp16, channels-last, eagerly and replayed from a CUDA graph.  This is a *reported
baseline* beside the hand-written kernels (bench.py key `library_baseline`), never part of the product.

Only the ops of the two graphs are interpreted (Conv, Sigmoid, Mul, Add, Concat, Split, MaxPool, Resize
nearest x2, Reshape, Transpose, Softmax, Slice, Sub, Div and the constant-folding shape ops).
Usage: python tools/library_baseline.py [armor_batch] [iters]   -> one JSON line
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rm_radar_b200 import onnx_wire  # noqa: E402


class CudnnNet:
    def __init__(self, path, device, dtype=torch.float16):
        self.g = onnx_wire.load(path)
        self.dev, self.dtype = device, dtype
        self.consts = {}
        for k, v in self.g.initializers.items():
            t = torch.from_numpy(np.array(v, copy=True))
            if t.is_floating_point():
                t = t.to(device=device, dtype=dtype)
                if t.dim() == 4:
                    t = t.contiguous(memory_format=torch.channels_last)
            self.consts[k] = t
        # arithmetic operands live on the device from the start (integer constants too): nothing is copied per call,
        # which is also what makes the forward pass capturable into a CUDA graph
        for n in self.g.nodes:
            if n.op in ("Mul", "Add", "Sub", "Div", "Concat"):
                for i in n.inputs:
                    if i in self.consts and not self.consts[i].is_cuda and device.type == "cuda":
                        self.consts[i] = self.consts[i].to(device)
        self.inp = self.g.inputs[0].name
        self.out = self.g.outputs[0].name
        self.static = None
        self.dynamic = None
        self._int_cache = {}

    def _ints(self, t):
        """Shape operands as Python ints.  They are constants (folded on the first call); a constant that also feeds
        device arithmetic lives on the device, so its values are read back once and remembered — nothing is copied
        on later calls, CUDA-graph capture included.  The tensor is kept with its values so its id stays unique."""
        if not isinstance(t, torch.Tensor):
            return [int(v) for v in t]
        hit = self._int_cache.get(id(t))
        if hit is None:
            hit = (t, [int(v) for v in t.tolist()])
            self._int_cache[id(t)] = hit
        return hit[1]

    @torch.no_grad()
    def __call__(self, x):
        """The first call also folds every node that does not depend on the input (anchor grids, shape arithmetic) and
        keeps its result on the device, so later calls — and the CUDA-graph capture — run device work only."""
        first = self.static is None
        if first:
            dyn = {self.inp}
            for n in self.g.nodes:
                if n.op != "Shape" and any(i in dyn for i in n.inputs):   # shapes are fixed for a fixed input size
                    dyn.update(n.outputs)
            self.dynamic = dyn
            self.static = {}
        env = dict(self.consts)
        env.update(self.static)
        env[self.inp] = x
        env[""] = None
        for n in self.g.nodes:
            if not first and not any(o in self.dynamic for o in n.outputs):
                continue
            ins = [env[i] for i in n.inputs]
            outs = self._run(n, ins)
            if not isinstance(outs, (tuple, list)):
                outs = (outs,)
            for name, val in zip(n.outputs, outs):
                if first and name not in self.dynamic and isinstance(val, torch.Tensor) and val.is_floating_point():
                    val = val.to(self.dev)
                env[name] = val
                if first and name not in self.dynamic:
                    self.static[name] = val
        if first and self.dev.type == "cuda":
            # folded operands of device arithmetic move to the device once (integer ones too): no copies per call
            for n in self.g.nodes:
                if n.op in ("Mul", "Add", "Sub", "Div", "Concat") and any(o in self.dynamic for o in n.outputs):
                    for i in n.inputs:
                        if i in self.static and isinstance(self.static[i], torch.Tensor) and not self.static[i].is_cuda:
                            self.static[i] = self.static[i].to(self.dev)
        return env[self.out]

    def _same_device(self, ins):
        """constant-folded shape arithmetic stays on the host; an operand that meets a device tensor follows it"""
        dev = [t for t in ins if isinstance(t, torch.Tensor) and t.is_cuda]
        if not dev:
            return ins
        return [t.to(dev[0].device) if isinstance(t, torch.Tensor) and not t.is_cuda else t for t in ins]

    def _run(self, n, ins):
        op, a = n.op, n.attrs
        if op in ("Mul", "Add", "Sub", "Div", "Concat"):
            ins = self._same_device(ins)
        if op == "Conv":
            p = a["pads"]
            return F.conv2d(ins[0], ins[1], ins[2] if len(ins) > 2 else None, stride=a["strides"], padding=(p[0], p[1]),
                            dilation=a["dilations"])
        if op == "Sigmoid":
            return torch.sigmoid(ins[0])
        if op == "Mul":
            return ins[0] * ins[1]
        if op == "Add":
            return ins[0] + ins[1]
        if op == "Sub":
            return ins[0] - ins[1]
        if op == "Div":
            if not ins[0].is_floating_point() and not ins[1].is_floating_point():
                return torch.div(ins[0], ins[1], rounding_mode="trunc")
            return ins[0] / ins[1]
        if op == "Concat":
            dev = [t for t in ins if t.is_cuda]
            if dev:
                ins = [t.to(dev[0].device) for t in ins]
            return torch.cat(ins, dim=a["axis"])
        if op == "Split":
            return torch.split(ins[0], self._ints(ins[1]), dim=a["axis"])
        if op == "MaxPool":
            p = a["pads"]
            return F.max_pool2d(ins[0], a["kernel_shape"], a["strides"], (p[0], p[1]))
        if op == "Resize":
            return F.interpolate(ins[0], scale_factor=2.0, mode="nearest")
        if op == "Reshape":
            shape = self._ints(ins[1])
            src = ins[0]
            shape = [src.shape[i] if s == 0 else s for i, s in enumerate(shape)]
            return src.reshape(shape)
        if op == "Transpose":
            return ins[0].permute(a["perm"])
        if op == "Softmax":
            return torch.softmax(ins[0], dim=a["axis"])
        if op == "Slice":
            data, starts, ends = ins[0], self._ints(ins[1]), self._ints(ins[2])
            axes = self._ints(ins[3]) if len(ins) > 3 and ins[3] is not None else list(range(len(starts)))
            steps = self._ints(ins[4]) if len(ins) > 4 and ins[4] is not None else [1] * len(starts)
            idx = [slice(None)] * data.dim()
            for s, e, ax, st in zip(starts, ends, axes, steps):
                dim = data.shape[ax]
                s = max(min(s + dim if s < 0 else s, dim), 0)
                e = max(min(e + dim if e < 0 else e, dim), 0)
                idx[ax] = slice(s, e, st)
            return data[tuple(idx)]
        if op == "Shape":
            return torch.tensor(list(ins[0].shape), dtype=torch.int64)
        if op == "Gather":
            ax = a.get("axis", 0)
            idx = ins[1]
            if idx.dim() == 0:
                return ins[0].select(ax, int(idx))
            return torch.index_select(ins[0], ax, idx.reshape(-1).to(ins[0].device)).reshape(
                list(ins[0].shape[:ax]) + list(idx.shape) + list(ins[0].shape[ax + 1:]))
        if op == "Unsqueeze":
            out = ins[0]
            for ax in sorted(self._ints(ins[1])):
                out = out.unsqueeze(ax)
            return out
        if op == "Cast":
            to = {1: self.dtype, 7: torch.int64, 6: torch.int32, 9: torch.bool}[a["to"]]
            return ins[0].to(to)
        if op == "Range":
            return torch.arange(ins[0].item(), ins[1].item(), ins[2].item(), dtype=ins[0].dtype)
        if op == "Expand":
            shape = self._ints(ins[1])
            return ins[0].expand(torch.broadcast_shapes(tuple(ins[0].shape), tuple(shape))).clone()
        if op == "ConstantOfShape":
            shape = self._ints(ins[0])
            v = a.get("value")
            val = v.array.reshape(-1)[0] if v is not None else 0.0
            dt = torch.from_numpy(np.asarray(val)).dtype if v is not None else torch.float32
            if dt.is_floating_point:
                dt = self.dtype
            return torch.full(shape, val.item() if hasattr(val, "item") else val, dtype=dt)
        if op == "Constant":
            t = torch.from_numpy(a["value"].array.copy())
            return t.to(device=self.dev, dtype=self.dtype) if t.is_floating_point() and t.numel() > 16 else t
        raise NotImplementedError(op)


def time_net(net, x, iters):
    """(eager ms, CUDA-graph replay ms) per forward, CUDA events on the current stream."""
    for _ in range(3):
        y = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = net(x)
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / iters
    graphed = None
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                y = net(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = net(x)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graphed = e0.elapsed_time(e1) / iters
    except Exception as e:   # noqa: BLE001  (host-side shape ops inside capture)
        graphed = None
        print("graph capture failed:", e, file=sys.stderr)
    return eager, graphed, y


def measure(engine_dir, armor_batch=7, iters=20, device=0):
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", device)
    out = {"kind": "torch + cuDNN fp16 channels-last (TensorRT stand-in, same ONNX graphs, same GPU)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    tot_e = tot_g = 0.0
    for name, b in (("car", 1), ("armor", armor_batch)):
        net = CudnnNet(os.path.join(engine_dir, name + ".onnx"), dev)
        x = torch.rand(b, 3, 640, 640, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        eager, graphed, y = time_net(net, x, iters)
        out[name] = {"batch": b, "eager_ms": eager, "graph_ms": graphed, "out_shape": list(y.shape)}
        tot_e += eager
        tot_g += graphed if graphed is not None else eager
    out["conv_stack_ms_eager"] = tot_e
    out["conv_stack_ms_graph"] = tot_g
    return out


if __name__ == "__main__":
    kb = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    it = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    print(json.dumps(measure(os.path.join(ROOT, "rm_radar_b200", "engines"), kb, it)))
