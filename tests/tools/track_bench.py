#!/usr/bin/env python
"""Tracker measurement (host code, no GPU needed): Tracker::update through the C ABI on the record
array `rmr_run_once` fills, against the numpy oracle.  python tests/tools/track_bench.py [robots] [frames]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rm_radar_b200 as rr  # noqa: E402
from oracle import track_oracle as to  # noqa: E402
from rm_radar_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
rng = np.random.default_rng(0)
pos = rng.uniform(-5, 5, (n, 3))
vel = rng.uniform(-1, 1, (n, 3))
trk, ora = rr.Tracker([0.2, 0.2, 0.2], 12), to.Tracker([0.2, 0.2, 0.2], 12)
recs = (_lib.RobotRec * n)()


def fill():
    for i in range(n):
        r = recs[i]
        r.is_detected, r.label, r.confidence, r.n_armors = 1, i % 12, 0.9, 1
        r.armors[0] = _lib.Detection(0, 0, 8, 8, float(i % 12), 0.9)
        r.is_located = 1
        r.location = (C.c_float * 3)(*pos[i])


t = 0
spent = 0.0
for k in range(frames):
    t += 33_000_000
    pos += vel * 0.033
    fill()
    t0 = time.perf_counter()
    trk.update_records(recs, n, t)
    spent += time.perf_counter() - t0
product_us = spent / frames * 1e6
m = max(frames // 100, 20)
t2, spent = 0, 0.0
for k in range(m):
    t2 += 33_000_000
    obs = [to.RobotObs(armors=[(i % 12, 0.9)], location=pos[i], label=i % 12) for i in range(n)]
    t0 = time.perf_counter()
    ora.update(obs, t2)
    spent += time.perf_counter() - t0
print(json.dumps({"robots": n, "tracks": len(trk.tracks()), "frames": frames,
                  "tracker_update_us_per_frame": product_us, "note": "C ABI call from Python included, one host core",
                  "oracle_numpy_us_per_frame": spent / m * 1e6, "host": os.uname().machine, "cores": os.cpu_count()}))
