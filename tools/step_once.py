#!/usr/bin/env python
"""A few `rmr_run_once` steps of the bench workload (BASELINE config[1]) with nothing around them:
the command ncu wraps for the launch list and the `--set full` captures of the non-conv kernels.
Usage: python tools/step_once.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import rm_radar_b200 as rr  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frame, bg, cloud, fx = bench.make_inputs(1)
dev = torch.device("cuda", 0)
W, H, NPTS = bench.W, bench.H, bench.NPTS
det = rr.RobotDetector(fx.engine("car"), fx.engine("armor"), (W, H), fx.CLASS_NUM, fx.MAX_BATCH, fx.OPT_BATCH, device=0)
loc = rr.Locator(W, H, fx.scaled_intrinsic(W, H), fx.LIDAR_TO_CAMERA, fx.WORLD_TO_CAMERA, device=0)
loc.update(bg[: 1 << 20])
f = torch.from_numpy(frame).to(dev)
c = torch.from_numpy(cloud).to(dev)
torch.cuda.synchronize()
for i in range(steps):
    recs, n = rr.run_once_records(det, loc, f.data_ptr(), True, W, H, W * 3, c.data_ptr(), True, NPTS, 12)
torch.cuda.synchronize()
print("robots", n, file=sys.stderr)
