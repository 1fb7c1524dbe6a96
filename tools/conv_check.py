#!/usr/bin/env python
"""Parity + timing of the tcgen05 conv on every distinct conv shape of car.onnx (batch 1) and armor.onnx
(batch B) through the C-ABI self-test hook (tcgen05 path vs the CUDA-core checker, random data).
Usage: python tools/conv_check.py [armor_batch] [iters]      (environment variables select kernel variants)"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rm_radar_b200 as rr  # noqa: E402

kb = int(sys.argv[1]) if len(sys.argv) > 1 else 7
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(r"conv\s+umma=1 in\s*(\d+)x\s*(\d+)x\s*(\d+) out\s*(\d+)x\s*(\d+)x\s*(\d+) k(\d) s(\d)")
shapes = []
batch = 1
for line in open(os.path.join(root, "profiles", "r1_layers.txt")):
    if line.startswith("== car"):
        batch = 1
    elif line.startswith("== armor"):
        batch = kb
    m = pat.search(line)
    if m:
        h, w, cin, ho, wo, cout, k, s = (int(v) for v in m.groups())
        shapes.append((batch, h, w, cin, cout, k, s))
count = {}
for sh in shapes:
    count[sh] = count.get(sh, 0) + 1
tot_ms = {1: 0.0, kb: 0.0}
tot_fl = {1: 0.0, kb: 0.0}
bad = 0
for sh, c in count.items():
    n, h, w, cin, cout, k, s = sh
    res = 1 if (k == 3 and s == 1 and cin == cout) else 0
    f32 = 1 if cout in (1, 12) or (k == 1 and cout == 64 and cin == 64) else 0
    act = 0 if f32 else 1
    d, ref, ms = rr.conv_selftest(n, h, w, cin, cout, k, s, act, res, f32, seed=1, iters=iters)
    tol = (2e-3 if f32 else 6e-3) * max(1.0, ref)
    ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
    fl = 2.0 * n * ho * wo * cout * k * k * cin
    ok = d == d and d <= tol
    bad += 0 if ok else 1
    tot_ms[n] += ms * c
    tot_fl[n] += fl * c
    print(f"n{n} {h:3d}x{w:3d} c{cin:4d}->{cout:4d} k{k}s{s} x{c}  diff {d:.2e} (tol {tol:.1e}) {'ok ' if ok else 'BAD'} "
          f"{ms * 1e3:8.2f} us {fl / ms / 1e9 if ms > 0 else 0:7.1f} TFLOP/s", flush=True)
for n in tot_ms:
    if tot_ms[n] > 0:
        print(f"batch {n}: sum over layers {tot_ms[n]:.4f} ms, {tot_fl[n] / 1e9:.2f} GFLOP, {tot_fl[n] / tot_ms[n] / 1e9:.1f} TFLOP/s; bad {bad}")
