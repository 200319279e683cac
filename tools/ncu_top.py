#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu -i rep --page source --csv | python tools/ncu_top.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
hdr = rows[1]
si = hdr.index("# Samples")
body = rows[2:]
tot = sum(int(r[si]) for r in body if r[si].isdigit())
order = sorted(range(len(body)), key=lambda i: -int(body[i][si]) if body[i][si].isdigit() else 0)[:n]
print(f"total samples {tot}, {len(body)} SASS instructions")
for i in sorted(order):
    r = body[i]
    print(f"{i:5d} {int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  {r[1].strip()}")
