#!/usr/bin/env python
"""Digest ncu output into the small files kept under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.md> [first_kernel]   per-kernel totals / shares of one bench step
                                                                    (first_kernel: keep from its last launch on = the last whole step)
  python tools/ncu_summary.py full <raw.csv> <out.json> <out.md>    key metrics of an `ncu --set full` capture
                                                                    (raw.csv = `ncu -i rep --page raw --csv`)
"""
import collections
import csv
import json
import sys


def launches(path, out_md, first_kernel=None):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    rows = list(r)
    if first_kernel:   # keep the last complete step: from the last launch of `first_kernel` to the end
        starts = [i for i, row in enumerate(rows) if first_kernel in row[ki]]
        if starts:
            rows = rows[starts[-1]:]
    for row in rows:
        name = row[ki].split("(")[0].replace("rmr::<unnamed>::", "").replace("void ", "")
        tot[name] += float(row[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    with open(out_md, "w") as f:
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in tot.most_common():
            f.write(f"| `{k[:60]}` | {cnt[k]} | {v / 1e3:.1f} | {v / cnt[k] / 1e3:.2f} | {100 * v / total:.1f} % |\n")
        f.write(f"| **all** | {sum(cnt.values())} | {total / 1e3:.1f} | | 100 % |\n")
    print(open(out_md).read())


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def full(path, out_json, out_md):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    recs = []
    for row in data:
        rec = {}
        for w, i in cols:
            v = row[i]
            try:
                v = float(v.replace(",", ""))
                if units[i] in UNIT and "bytes" in w:
                    v *= UNIT[units[i]]
            except ValueError:
                pass
            rec[w] = v
        recs.append(rec)
    dram = [r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0) for r in recs]
    summary = {"kernels": len(recs), "dram_bytes_per_launch_mean": sum(dram) / max(len(dram), 1),
               "tensor_pipe_pct_mean": sum(r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for r in recs) / max(len(recs), 1),
               "launches": recs}
    json.dump(summary, open(out_json, "w"), indent=1)
    with open(out_md, "w") as f:
        f.write("| grid | time us | regs | tensor pipe % | SM % | DRAM % | dram MB | L2 MB |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in recs:
            f.write("| {} | {:.2f} | {:.0f} | {:.1f} | {:.1f} | {:.1f} | {:.2f} | {:.2f} |\n".format(
                r.get("Grid Size"), r.get("gpu__time_duration.sum", 0), r.get("launch__registers_per_thread", 0),
                r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0),
                r.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0),
                r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0),
                (r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e6, r.get("lts__t_sectors.sum", 0) * 32 / 1e6))
    print(open(out_md).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4])
