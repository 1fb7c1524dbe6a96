#!/usr/bin/env python
"""HBM-side accounting of the non-conv kernels whose algorithmic bytes follow exactly from the workload definition:
bytes per launch divided by the launch durations in the committed ncu launch list (profiles/r1_launches_step.csv:
gpu__time_duration, cold cache, serialised) against the measured HBM peak.
    python tools/aux_roofline.py > profiles/r1_aux_kernels.md
Workload of the list: BASELINE configs[1] -- 7 armour ROIs, 640x640 network input, 100 000-point cloud."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6550.7
CAR_ANCHORS = 160 * 160 + 80 * 80 + 40 * 40 + 20 * 20        # P2 head, strides 4..32
ARMOR_ANCHORS = 80 * 80 + 40 * 40 + 20 * 20 + 10 * 10        # P6 head, strides 8..64
KERNELS = {   # name -> {grid: (what, bytes)}
    "decode_compact_kernel": {
        "(133, 1, 1)": ("car head: 34 000 anchors x (64 DFL bins + 1 class) fp32 read once", CAR_ANCHORS * 65 * 4),
        "(34, 7, 1)": ("armour head, 7 ROIs: 8 500 anchors x (64 + 12) fp32 read once", 7 * ARMOR_ANCHORS * 76 * 4),
    },
    "project_kernel": {"(391, 1, 1)": ("100 000 points: 12 B read + one 64-bit atomicMax each", 100_000 * 20)},
}
rows = [l for l in open(os.path.join(ROOT, "profiles", "r1_launches_step.csv")) if l.startswith('"')]
r = csv.reader(rows)
hdr = next(r)
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
print("| kernel | grid | µs (ncu, cold) | algorithmic bytes | what they are | GB/s | of measured HBM peak |")
print("|---|---|---:|---:|---|---:|---:|")
for row in r:
    name = row[ki].split("(")[0].split("::")[-1]
    spec = KERNELS.get(name, {}).get(row[gi])
    if spec is None:
        continue
    us = float(row[vi].replace(",", "")) / 1e3
    gbs = spec[1] / us / 1e3
    print(f"| `{name}` | {row[gi]} | {us:.2f} | {spec[1] / 1e6:.2f} MB | {spec[0]} | {gbs:.0f} | {100 * gbs / peak:.1f} % |")
print(f"\nMeasured HBM peak: {peak:.0f} GB/s (MEASURED_PEAKS.json); 6 MB at that rate is 1 µs, so kernels of this size sit in the launch / "
      "first-touch latency regime.  The other non-conv kernels (letterbox, NMS, locate) move less than that per launch; their "
      "durations are in r1_launches_step.md.  An `ncu --set full` capture of these kernels (DRAM bytes instead of algorithmic "
      "bytes) was not taken in round 1: the attempt ran into the end of the GPU budget.")
