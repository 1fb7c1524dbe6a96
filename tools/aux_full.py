#!/usr/bin/env python
"""Digest of the `ncu --set full` capture of the non-conv kernels of one bench step (BASELINE config[1]):
DRAM bytes beside the algorithmic bytes, achieved GB/s against the measured HBM peak.
Usage: python tools/aux_full.py gpurun_out/r2_aux_raw.csv profiles/r2_aux_full.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
W, H, NPTS, K = 1920, 1080, 100_000, 7
WZ, HZ = 1920 * 0 + 720, 0   # filled below from the launch geometry when needed

# algorithmic bytes per launch (what the kernel must move once), keyed by (kernel, grid.y or grid.x)
def algo_bytes(name, grid):
    gx, gy, _ = grid
    if name == "letterbox_fused_kernel":      # whole frame -> 640x640x4 fp16
        return W * H * 3 + 640 * 640 * 4 * 2, "frame u8 read once + fp16 NHWC4 blob written"
    if name == "letterbox_stage_kernel":      # K ROI crops (bilinear footprint ~ crop area) -> u8 staging
        return None, "K ROI crops of the resident frame -> u8 staging (crop sizes vary)"
    if name == "letterbox_blob_kernel":
        return gy * (640 * 640 * 3 + 640 * 640 * 4 * 2), "u8 staging read + fp16 NHWC4 blob written, K images"
    if name == "conv_stem_kernel":
        b = 1 if gx < 1000 else K
        return b * (640 * 640 * 4 * 2 + 320 * 320 * 32 * 2), "fp16 NHWC4 input + 32-channel fp16 output"
    if name == "decode_compact_kernel":
        return (gy * 8500 * (64 + 12) * 4) if gy > 1 else 34000 * (64 + 1) * 4, "head logits (64 DFL bins + classes) fp32, read once"
    if name == "project_kernel":
        return NPTS * 12 + NPTS * 8, "12 B per point read + one 64-bit atomic per point"
    if name == "copy_channels_kernel":
        return None, "Concat copy (read + write of the view)"
    return None, ""


def main(raw, out_md):
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, k, scale_bytes=False):
        v = float(r[ix[k]].replace(",", ""))
        if scale_bytes:
            v *= UNIT.get(units[ix[k]], 1.0)
        return v

    with open(out_md, "w") as f:
        f.write("# Non-conv kernels of one bench step (BASELINE config[1], K = 7 ROIs): `ncu --set full --clock-control none`\n\n")
        f.write(f"Measured HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json). `dram` = dram__bytes_read.sum + dram__bytes_write.sum of the launch; "
                "`algo` = bytes the kernel must move once; GB/s = max(dram, algo) / duration.  ncu durations are cold-cache and serialised.\n\n")
        f.write("| kernel | grid | us | regs | dram MB | algo MB | dram/algo | L2 MB | GB/s | of peak | SM % | what moves |\n")
        f.write("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        seen = set()
        for r in data:
            name = r[ix["Kernel Name"]].split("(")[0].replace("unnamed>::", "").replace("void ", "").split("<")[0].strip()
            grid = tuple(int(v) for v in r[ix["Grid Size"]].strip("()").split(","))
            key = (name, grid)
            if key in seen:
                continue
            seen.add(key)
            us = val(r, "gpu__time_duration.sum")
            dram = val(r, "dram__bytes_read.sum", True) + val(r, "dram__bytes_write.sum", True)
            l2 = val(r, "lts__t_sectors.sum") * 32 if "lts__t_sectors.sum" in ix else 0
            regs = val(r, "launch__registers_per_thread")
            smp = val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
            ab, what = algo_bytes(name, grid)
            moved = max(dram, ab or 0)
            gbs = moved / (us * 1e-6) / 1e9
            f.write(f"| `{name}` | {grid} | {us:.2f} | {regs:.0f} | {dram / 1e6:.2f} | {(ab / 1e6) if ab else float('nan'):.2f} | "
                    f"{(dram / ab) if ab else float('nan'):.2f} | {l2 / 1e6:.2f} | {gbs:.0f} | {100 * gbs / peak:.1f} % | {smp:.1f} | {what} |\n")
    print(open(out_md).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
