#!/usr/bin/env python
"""PCD ingestion: device parser vs the oracle's reader on a synthetic 1 M-point ASCII file
shaped like the reference's frame clouds (integer millimetres, `x y z` per line).  Wall-clock around the
C ABI call (pinned staging copy + H2D + 4 kernels + Locator::update's projection), median of 10."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rm_radar_b200 as rr  # noqa: E402
from oracle import locate_oracle as lo  # noqa: E402
from tests import fixtures as fx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
pts = np.stack([rng.integers(5000, 29000, n), rng.integers(-8000, 8000, n), rng.integers(-2000, 3000, n)], 1)
hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
       f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ascii\n")
blob = hdr.encode() + "\n".join("%d %d %d" % tuple(p) for p in pts).encode() + b"\n"
path = "/tmp/pcd_bench.pcd"
open(path, "wb").write(blob)

loc = rr.Locator(3840, 2160, fx.scaled_intrinsic(3840, 2160), fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA)
got = rr.pcd_parse(blob)
assert np.array_equal(got, pts.astype(np.float32)), "device parse differs"
ts = []
for _ in range(12):
    t0 = time.perf_counter()
    loc.update_pcd(blob)
    loc.stats()                      # synchronises the locator stream
    ts.append(time.perf_counter() - t0)
gpu = float(np.median(ts[2:]))
t0 = time.perf_counter()
ref = lo.read_pcd(path)
cpu = time.perf_counter() - t0
assert np.array_equal(ref, got)
print(f"{n} points, {len(blob) / 1e6:.1f} MB ascii: device ingest+update {gpu * 1e3:.2f} ms "
      f"({len(blob) / gpu / 1e9:.1f} GB/s of text incl. the PCIe upload); oracle reader (numpy, 1 core) {cpu * 1e3:.0f} ms "
      f"({len(blob) / cpu / 1e6:.0f} MB/s); speed-up {cpu / gpu:.0f}x")
