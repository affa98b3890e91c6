#!/usr/bin/env python
"""Per source line: executed warp instructions and stall samples of one kernel from an ncu report.
usage: ncu_lines.py <report.ncu-rep> <kernel name> [top N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ii, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
data = []
seen_first_kernel = False
for r in rows[hi + 1:]:
    if r and r[0] == "Line No":
        break  # a second launch of the same kernel
    if r and r[0] and r[0].isdigit():
        try:
            data.append((int(r[ii]), int(r[si]), int(r[ti]), int(r[0]), r[1]))
        except ValueError:
            pass
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("kernel %s: %d warp instructions, %d samples" % (kern, tot, ts))
for d in sorted(data, key=lambda x: -x[0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  thr/inst %4.1f  L%-5d %s" % (100 * d[0] / tot, 100 * d[1] / max(1, ts), d[2] / max(1, d[0]), d[3], d[4].strip()[:105]))
