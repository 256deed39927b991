#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: warp instructions per opcode and per source line
of the first kernel section.  Usage: ncu_opcodes.py file.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = None
data = []
for r in rows:
    if len(r) > 5 and r[0] == "Address":
        if hdr is not None:
            break
        hdr = r
        continue
    if hdr is not None and len(r) == len(hdr):
        data.append(r)
iI, iS, iSamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[iI]) for r in data)
print("total warp instr %d, sass lines %d" % (tot, len(data)))
c, cs = Counter(), Counter()
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[iI])
    cs[op] += int(r[iSamp])
ts = sum(cs.values())
for op, n in c.most_common(top):
    print("%-10s %10d %5.1f%%   stall samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * cs[op] / max(ts, 1)))
