#!/usr/bin/env python
"""Per-k-block stamps of the first tile of a conv2.cu CTA (RMR_DBG_MODE=1): A issue, B issue, operands consumed."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RMR_DBG_MODE"] = "1"
import rm_radar_b200 as rr  # noqa: E402

for a in sys.argv[1:]:
    sh = tuple(int(v) for v in a.split(','))
    t = rr.conv_timeline(*sh)
    d = t - t[:, 0:1]
    for c in (0, len(t) // 2):
        r = d[c]
        print(f"== {sh} cta {c}: setup {r[1]}")
        print("   A issue  ", [int(v) for v, raw in zip(r[4:24], t[c, 4:24]) if raw])
        print("   pre-wait ", [int(v) for v, raw in zip(r[24:44], t[c, 24:44]) if raw])
        print("   consumed ", [int(v) for v, raw in zip(r[44:60], t[c, 44:60]) if raw])
