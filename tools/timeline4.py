#!/usr/bin/env python
"""Epilogue steps of warp 4 per tile (RMR_DBG_MODE=2), SM cycles since CTA start:
loop top | accumulator seen | first chunk in registers | activation done (before the staging wait) | staging free |
staged (STS done) | fence + arrive done | DMA warp saw sub-tile 0 ready."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RMR_DBG_MODE"] = "2"
import rm_radar_b200 as rr  # noqa: E402

names = ["top", "acc", "ld", "act", "free", "sts", "arrive", "dma"]
for a in sys.argv[1:]:
    sh = tuple(int(v) for v in a.split(','))
    t = rr.conv_timeline(*sh)
    d = t - t[:, 0:1]
    for c in (0, len(t) // 2):
        print(f"== {sh} cta {c}: setup {d[c, 1]}")
        for ti in range(7):
            if t[c, 3 + 8 * ti] == 0:
                break
            row = d[c, 2 + 8 * ti: 10 + 8 * ti]
            print("   tile %d: " % ti + "  ".join(f"{nm} {int(v) if t[c, 2 + 8 * ti + i] else -1}" for i, (nm, v) in enumerate(zip(names, row))))
