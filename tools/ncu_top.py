#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`
(one section per profiled kernel).  usage: ncu -i rep --page source --csv | python tools/ncu_top.py [N]"""
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "body": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["body"].append(r)
for s in sections:
    hdr, body = s["hdr"], s["body"]
    si = hdr.index("# Samples")
    tot = sum(int(r[si]) for r in body if r[si].isdigit())
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si]) if body[i][si].isdigit() else 0)[:n]
    print(f"== {s['name'][:60]}: total samples {tot}, {len(body)} SASS instructions")
    for i in sorted(order):
        r = body[i]
        print(f"{i:5d} {int(r[si]):6d} {100 * int(r[si]) / max(tot, 1):5.1f}%  {r[1].strip()}")
