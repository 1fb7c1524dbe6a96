#!/usr/bin/env python
"""Prints the per-CTA phase timeline (SM cycles) of the tcgen05 conv kernel for a few layer shapes."""
import numpy as np
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rm_radar_b200 as rr  # noqa: E402

import sys
SHAPES = [(1, 160, 160, 64, 64, 3, 1), (1, 160, 160, 64, 1, 1, 1), (1, 20, 20, 512, 64, 3, 1), (1, 40, 40, 128, 128, 3, 1),
          (7, 80, 80, 128, 128, 3, 1)]
if len(sys.argv) > 1:
    SHAPES = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for sh in SHAPES:
    t = rr.conv_timeline(*sh)
    n = len(t)
    t0 = t[:, 0:1]
    d = t - t0
    print(f"== shape {sh}: {n} CTAs; kernel span {int((t[:, 21].max() - t[:, 0].min()))} cycles (cross-SM clocks, approximate)")
    for c in sorted(set([0, 1, n // 2, n - 1])):
        r = d[c]
        nit = int(np.sum(t[c, 2:18] != 0))
        print(f" cta {c}: setup {r[1]}, full[it] {[int(x) for x in r[2:2 + nit]]}, mma_issued {r[18]}, acc_ready {r[19]}, "
              f"epi_done {r[20]}, exit {r[21]}; tma_issue[it] {[int(x) for x in r[24:24 + nit]]}; epi chunk0: ld_done {r[40]} stored {r[42]}; mma pre-wait {[int(x) for x in r[44:44 + min(nit, 10)]]}; B issue {[int(x) for x in r[54:54 + min(nit, 10)]]}")
    med = np.median(d, axis=0)
    print(f" median: setup {med[1]:.0f} first_full {med[2]:.0f} mma_issued {med[18]:.0f} acc_ready {med[19]:.0f} epi_done {med[20]:.0f} exit {med[21]:.0f}")
