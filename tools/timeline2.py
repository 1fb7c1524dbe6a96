#!/usr/bin/env python
"""Per-CTA phase timeline (SM cycles) of the round-2 tcgen05 conv kernel (conv2.cu) for a few layer shapes.
Slot map: 0 start, 1 setup done, 2 + 8 t + {0 A producer past its empty wait, 1 issuer past the accumulator wait,
2 first operands seen, 3 all MMAs issued, 4 epilogue sees the accumulator, 5 accumulator read out, 6 / 7 first /
second epilogue group stored} for tiles t < 7, 60 exit.
Usage: RMR_CONV_V2=1 python tools/timeline2.py [n,h,w,cin,cout,k,stride ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rm_radar_b200 as rr  # noqa: E402

SHAPES = [(7, 160, 160, 64, 64, 1, 1), (7, 160, 160, 32, 32, 3, 1), (7, 80, 80, 64, 64, 3, 1), (7, 80, 80, 128, 128, 3, 1),
          (7, 40, 40, 128, 128, 3, 1), (7, 80, 80, 128, 256, 3, 2), (1, 20, 20, 256, 256, 3, 1), (1, 80, 80, 128, 128, 1, 1)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
names = ["A>empty", "acc-free", "operands", "issued", "acc-full", "read-out", "stored0", "stored1"]
for sh in SHAPES:
    t = rr.conv_timeline(*sh)
    n = len(t)
    d = t - t[:, 0:1]
    print(f"== shape {sh}: {n} CTAs")
    for c in sorted(set([0, n // 2, n - 1])):
        r = d[c]
        print(f" cta {c}: setup {r[1]}  exit {r[60]}")
        for ti in range(7):
            row = r[2 + 8 * ti: 10 + 8 * ti]
            if t[c, 3 + 8 * ti] == 0:
                break
            print("   tile %d: " % ti + "  ".join(f"{nm} {int(v) if t[c, 2 + 8 * ti + i] else -1}" for i, (nm, v) in enumerate(zip(names, row))))
    med = np.median(d, axis=0)
    per = [(med[5 + 8 * (ti + 1)] - med[5 + 8 * ti]) for ti in range(5) if np.median(t[:, 5 + 8 * (ti + 1)]) > 0]
    print(f" median: setup {med[1]:.0f}, first operands {med[4]:.0f}, exit {med[60]:.0f}; issue period per tile {[int(x) for x in per]}")
