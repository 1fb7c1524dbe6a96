#!/usr/bin/env python
"""Algorithmic bytes per tcgen05 conv launch (activations in + out + residual + weights), from the engine plan.

  python tools/algo_bytes.py            # car at batch 1, armor at batch 7 (the bench workload)
The figure profiles/r1_summary.md §3 compares with ncu's dram__bytes per launch.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rm_radar_b200 import engine as E  # noqa: E402

for name, batch in (("car", 1), ("armor", 7)):
    plan = E.compile_onnx(os.path.join(ROOT, "rm_radar_b200", "engines", f"{name}.onnx"))
    act = wts = n = 0
    for op in plan.ops:
        if op.type != E.OP_CONV or op.src.C < 8:      # the 3-channel stem is a SIMT kernel
            continue
        out_sz = 2 if plan.bufs[op.dst.buf].dtype == E.DT_F16 else 4
        a = op.src.H * op.src.W * op.src.C * 2 + op.dst.H * op.dst.W * op.dst.C * out_sz
        if op.res is not None:
            a += op.res.H * op.res.W * op.res.C * 2
        act += batch * a
        wts += op.k * op.k * op.cin_pad * op.cout_pad * 2
        n += 1
    print(f"{name} (batch {batch}): {n} convs, mean {(act + wts) / n / 1e6:.2f} MB per launch "
          f"(activations {act / n / 1e6:.2f} + weights {wts / n / 1e6:.2f})")
