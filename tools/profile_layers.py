#!/usr/bin/env python
"""Per-layer table of the two conv stacks on the GPU: shape, path, ms, achieved TFLOP/s.
Usage: python tools/profile_layers.py [armor_batch] > gpurun_out/layers.txt"""
import sys

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rm_radar_b200 as rr  # noqa: E402
from tests import fixtures as fx

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 7
for name, classes, batch in (("car", 1, 1), ("armor", 12, kb)):
    det = rr.Detector(fx.engine(name), classes, (1920, 1080), max(batch, 1))
    rows = det.profile_ops(batch, 30)
    tot_ms = sum(r["ms"] for r in rows)
    tot_fl = sum(r["flops"] for r in rows)
    print(f"== {name} batch {batch}: {len(rows)} ops, sum of per-op ms {tot_ms:.4f}, {tot_fl / 1e9:.2f} GFLOP, "
          f"{tot_fl / tot_ms / 1e9:.1f} TFLOP/s if run back to back; graph replay {det.time_forward(batch, 30):.4f} ms")
    tn = {0: "conv", 1: "maxpool5", 2: "upsample2", 3: "copy"}
    for i, r in enumerate(rows):
        tf = r["flops"] / r["ms"] / 1e9 if r["ms"] > 0 else 0
        print(f"{i:3d} {tn[int(r['type'])]:9s} umma={int(r['umma'])} in {int(r['h_in']):3d}x{int(r['w_in']):3d}x{int(r['cin']):4d} "
              f"out {int(r['h_out']):3d}x{int(r['w_out']):3d}x{int(r['cout']):4d} k{int(r['k'])} s{int(r['stride'])} "
              f"{r['ms'] * 1e3:8.2f} us {tf:8.1f} TFLOP/s")
