"""ORACLE-side simulator (test infrastructure only): executes a compiled engine plan
(`rm_radar_b200.engine.Plan`) op by op in torch on the CPU, with the same buffer / channel-offset
semantics the CUDA runtime uses, plus a restatement of the fused DECODE tail.  It exists to check
the engine *compiler* (fusions, concat homes, split views, weight packing) and the decode-tail
algebra against the unmodified ONNX graph (oracle/onnx_torch.py) without a GPU."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from rm_radar_b200 import engine as E


def run_plan(plan: E.Plan, x_nchw: np.ndarray, half_weights: bool = True) -> list:
    """x_nchw: f32 [B,3,H,W].  Returns the per-level logits [B,H,W,64+nc] f32."""
    B = x_nchw.shape[0]
    bufs = [torch.zeros(B, b.H, b.W, b.C, dtype=torch.float32) for b in plan.bufs]
    bufs[plan.input_buf][..., :3] = torch.from_numpy(x_nchw).permute(0, 2, 3, 1)
    blob = bytes(plan.blob)
    for op in plan.ops:
        s, d = op.src, op.dst
        src = bufs[s.buf][..., s.coff:s.coff + s.C]
        if op.type == E.OP_CONV:
            kk = op.k * op.k
            w = np.frombuffer(blob, np.float16, op.cout_pad * kk * op.cin_pad, op.w_off)
            w = torch.from_numpy(w.astype(np.float32)).reshape(op.cout_pad, op.k, op.k, op.cin_pad)
            b = torch.from_numpy(np.frombuffer(blob, np.float32, op.cout_pad, op.b_off).copy())
            w = w[:d.C, :, :, :s.C].permute(0, 3, 1, 2)
            y = F.conv2d(src.permute(0, 3, 1, 2), w, b[:d.C], stride=op.stride, padding=op.k // 2)
            if op.act:
                y = y * torch.sigmoid(y)
            y = y.permute(0, 2, 3, 1)
            if op.res is not None:
                r = op.res
                y = y + bufs[r.buf][..., r.coff:r.coff + r.C]
        elif op.type == E.OP_MAXPOOL5:
            y = F.max_pool2d(src.permute(0, 3, 1, 2), 5, 1, 2).permute(0, 2, 3, 1)
        elif op.type == E.OP_UPSAMPLE2:
            y = src.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        elif op.type == E.OP_COPY:
            y = src
        else:
            raise NotImplementedError
        assert y.shape[1:3] == (d.H, d.W) and y.shape[3] == d.C, (op.name, y.shape, d)
        bufs[d.buf][..., d.coff:d.coff + d.C] = y
    return [bufs[b][..., :64 + plan.num_classes].numpy() for (b, H, W, s) in plan.levels]


def decode_tail(levels: list, plan: E.Plan) -> np.ndarray:
    """Restatement of the Ultralytics v8 Detect tail as exported in car.onnx / armor.onnx
    (SURVEY.md Appendix A 'Tail'): → [B, 4+nc, A] f32, same layout as the ONNX output."""
    outs = []
    for lv, (b, H, W, stride) in zip(levels, plan.levels):
        B = lv.shape[0]
        t = torch.from_numpy(lv).reshape(B, H * W, -1)
        box = t[..., :64].reshape(B, H * W, 4, 16)
        p = torch.softmax(box, dim=-1)
        dist = (p * torch.arange(16, dtype=torch.float32)).sum(-1)      # [B,A,4] l,t,r,b
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5,
                                torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")
        anc = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)          # [A,2]
        x1y1 = anc - dist[..., :2]
        x2y2 = anc + dist[..., 2:]
        cxy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        cls = torch.sigmoid(t[..., 64:])
        o = torch.cat([cxy * stride, wh * stride, cls], -1)             # [B,A,4+nc]
        outs.append(o)
    return torch.cat(outs, 1).permute(0, 2, 1).numpy()
