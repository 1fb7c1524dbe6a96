"""ORACLE (test infrastructure only — never imported by the product path).

fp32 torch interpreter for the ONNX graphs the reference feeds to TensorRT
(`/root/reference/src/detect/detector.h:122` enqueueV3; engine built from the sibling .onnx at
`/root/reference/src/detect/detector.cpp:177-243`).  TensorRT itself is a closed third-party
dependency absent from /root/reference and from this image, and the reference has no golden
vectors at the network boundary (`test/detect/detector_test.cpp:70-89` only checks a count), so
at this boundary parity is UNPINNED by the reference; the substitute pin is agreement with a
second independent engine, `cv2.dnn`, on the static car graph (tests/test_oracle_net.py).

Only the op set that occurs in car.onnx / armor.onnx / yolov8n.onnx is interpreted
(SURVEY.md Appendix C.2).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rm_radar_b200 import onnx_wire  # noqa: E402  (wire reader is shared host code, not compute)


class OnnxNet:
    def __init__(self, path: str, dtype=torch.float32):
        self.graph = onnx_wire.load(path)
        self.dtype = dtype
        self.consts = {}
        for k, v in self.graph.initializers.items():
            t = torch.from_numpy(np.array(v, copy=True))
            if t.is_floating_point():
                t = t.to(dtype)
            self.consts[k] = t
        self.input_name = self.graph.inputs[0].name
        self.output_name = self.graph.outputs[0].name

    @torch.no_grad()
    def __call__(self, x: np.ndarray | torch.Tensor, want=None):
        """Run the graph.  `want`: optional list of tensor names to also return."""
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        env = dict(self.consts)
        env[self.input_name] = x.to(self.dtype)
        env[""] = None
        for n in self.graph.nodes:
            ins = [env[i] for i in n.inputs]
            outs = self._run(n, ins)
            if not isinstance(outs, (tuple, list)):
                outs = (outs,)
            for name, val in zip(n.outputs, outs):
                env[name] = val
        out = env[self.output_name]
        if want is not None:
            return out, {k: env[k] for k in want}
        return out

    @staticmethod
    def _ints(t):
        return [int(v) for v in (t.tolist() if isinstance(t, torch.Tensor) else t)]

    def _run(self, n, ins):
        op, a = n.op, n.attrs
        if op == "Conv":
            w = ins[1]
            b = ins[2] if len(ins) > 2 else None
            p = a["pads"]
            assert p[0] == p[2] and p[1] == p[3] and a["group"] == 1
            return F.conv2d(ins[0], w, b, stride=a["strides"], padding=(p[0], p[1]),
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
            return torch.cat(ins, dim=a["axis"])
        if op == "Split":
            return torch.split(ins[0], self._ints(ins[1]), dim=a["axis"])
        if op == "MaxPool":
            p = a["pads"]
            assert a.get("ceil_mode", 0) == 0
            return F.max_pool2d(ins[0], a["kernel_shape"], a["strides"], (p[0], p[1]))
        if op == "Resize":
            assert a["mode"] == "nearest" and a["coordinate_transformation_mode"] == "asymmetric"
            scales = ins[2].tolist()
            assert scales[:2] == [1.0, 1.0] and scales[2] == scales[3] == 2.0
            return ins[0].repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
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
            return torch.index_select(ins[0], ax, idx.reshape(-1)).reshape(
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
            return torch.from_numpy(a["value"].array.copy())
        raise NotImplementedError(op)
