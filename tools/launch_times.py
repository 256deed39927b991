#!/usr/bin/env python3
"""Per-kernel durations (and DRAM bytes when present) from an `ncu --csv` launch list.
Usage: launch_times.py file.csv [substring ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
pats = sys.argv[2:]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i
        break
ik, iv, im, iid = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
d = {}
for r in rows[start + 2:]:
    if len(r) <= iv:
        continue
    d.setdefault((int(r[iid]), r[ik].split("(")[0][-48:]), {})[r[im]] = float(r[iv].replace(",", ""))
for (i, k), m in sorted(d.items()):
    if not pats or any(s in k for s in pats):
        t = m.get("gpu__time_duration.sum", 0.0) / 1e3
        rd, wr = m.get("dram__bytes_read.sum"), m.get("dram__bytes_write.sum")
        extra = "" if rd is None else "  dram rd %.1f MB wr %.1f MB  %.2f TB/s" % (rd / 1e6, wr / 1e6, (rd + wr) / t / 1e6)
        print("%4d %-48s %9.2f us%s" % (i, k, t, extra))
