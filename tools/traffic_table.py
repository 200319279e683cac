#!/usr/bin/env python3
"""DRAM traffic per kernel family from an ncu launch list captured with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
Writes a JSON (profiles/rNN_traffic.json) that bench.py reads for `roofline.traffic`."""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maf_yolo_b200 import engine as E  # noqa: E402
from maf_yolo_b200 import topology as T  # noqa: E402

src, out = sys.argv[1], sys.argv[2]
variant = sys.argv[3] if len(sys.argv) > 3 else "n"
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 32
rows = list(csv.reader(open(src)))
hdr, per_id = None, {}
for r in rows:
    if r and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        e = per_id.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})
        v = float(d["Metric Value"].replace(",", ""))
        unit = d["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
        e[d["Metric Name"]] = v * scale
launches = [per_id[k] for k in sorted(per_id)]
starts = [i for i, l in enumerate(launches) if "stem_conv" in l["name"]]
start = starts[min(1, len(starts) - 1)]
plan = E.Plan(T.build_graph(variant), 640, 640)
ours = [l for l in launches[start:] if "mafb200" in l["name"]]
fam = {}
for op, l in zip([o for o in plan.ops if o.kind != "detect_reset"], ours):
    f = fam.setdefault(op.kind, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0,
                                 "algorithmic_bytes": 0.0})
    f["launches"] += 1
    f["us"] += l.get("gpu__time_duration.sum", 0.0)
    f["dram_read_bytes"] += l.get("dram__bytes_read.sum", 0.0)
    f["dram_write_bytes"] += l.get("dram__bytes_write.sum", 0.0)
    f["algorithmic_bytes"] += op.bytes_per_image * batch
for f in fam.values():
    f["traffic_bytes_per_launch"] = (f["dram_read_bytes"] + f["dram_write_bytes"]) / f["launches"]
    f["algorithmic_bytes_per_launch"] = f["algorithmic_bytes"] / f["launches"]
json.dump({"variant": variant, "batch": batch, "source": os.path.basename(src), "families": fam}, open(out, "w"), indent=1)
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k:10s} n={f['launches']:3d} {f['us']:8.1f} us  dram {1e-6 * (f['dram_read_bytes'] + f['dram_write_bytes']):8.1f} MB"
          f"  algorithmic {1e-6 * f['algorithmic_bytes']:8.1f} MB")
