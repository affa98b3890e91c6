#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run in the build container, no GPU needed).

  ncu_summary.py launches <launches.csv> <out_prefix>    per-launch list (trimmed) + per-kernel sums
  ncu_summary.py full <report.ncu-rep> <out.csv>         one column per profiled launch, the metrics that matter
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]
UNIT = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "s": 1e3}


def clean(name):
    """kernel name without return type, template arguments and parameter list"""
    n = name.split("(")[0]
    if n.startswith("void "):
        n = n[5:]
    return n.split("<")[0]


def launches(path, prefix):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    tot, cnt = collections.defaultdict(float), collections.Counter()
    with open(prefix + "_launches.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_us"])
        for r in rows[hi + 1:]:
            if len(r) <= vi:
                continue
            ms = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
            name = clean(r[ki])
            tot[name] += ms
            cnt[name] += 1
            w.writerow([r[0], name, r[gi], r[bi], "%.3f" % (ms * 1e3)])
    T = sum(tot.values())
    with open(prefix + "_launch_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct", "avg_us"])
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            w.writerow([k, cnt[k], "%.3f" % v, "%.1f" % (100 * v / T), "%.1f" % (1e3 * v / cnt[k])])
        w.writerow(["total", sum(cnt.values()), "%.3f" % T, "100.0", ""])


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    names = []
    for r in rows[2:]:
        n = clean(r[kn])
        names.append(n + ("#%d" % (sum(1 for x in names if x.split("#")[0] == n) + 1) if any(x.split("#")[0] == n for x in names) else ""))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in rows[2:]])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
