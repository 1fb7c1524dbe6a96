#!/usr/bin/env python
"""JPEG stage measurement: device decode vs cv2.imdecode (the reference's cv::imread) on the
reference's own 2592x2048 frame.  python tools/jpeg_bench.py [iters]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rm_radar_b200 as rr  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
data = open(os.path.join(ROOT, "tests", "golden", "frames", "0.jpg"), "rb").read()
meta = rr.jpeg_info(data)
dec = rr.JpegDecoder(0)
stream = torch.cuda.Stream()
dec.set_stream(stream.cuda_stream)
for _ in range(5):
    dec.decode_device(data)
st = dec.status()
prof = [dec.profile(data) for _ in range(5)][-1]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0.record(stream)
for _ in range(iters):
    dec.decode_device(data)
e1.record(stream)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / iters * 1e3
dev = e0.elapsed_time(e1) / iters
line = {"file_bytes": len(data), "width": meta["width"], "height": meta["height"], "status": st,
        "device_ms_per_frame_incl_upload": dev, "wall_ms_per_frame": wall,
        "raw_frame_bytes": meta["width"] * meta["height"] * 3,
        "stage_ms": {k: round(v, 4) for k, v in prof.items()}}
try:
    import cv2
    buf = np.frombuffer(data, np.uint8)
    cv2.imdecode(buf, cv2.IMREAD_COLOR)
    t0 = time.perf_counter()
    n = max(iters // 5, 5)
    for _ in range(n):
        cv2.imdecode(buf, cv2.IMREAD_COLOR)
    line["cv2_imdecode_ms_per_frame"] = (time.perf_counter() - t0) / n * 1e3
    line["speedup_vs_cv2"] = line["cv2_imdecode_ms_per_frame"] / wall
except ImportError:
    pass
import json  # noqa: E402
print(json.dumps(line))
