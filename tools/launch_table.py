#!/usr/bin/env python3
"""Joins an `ncu --metrics gpu__time_duration.sum` launch list (CSV) of one eager forward with the
engine's plan: per-launch time, algorithmic GB/s and TFLOP/s, plus per-kernel-family totals.

  python tools/launch_table.py gpurun_out/launches.csv [--variant n] [--batch 32] [--forward 1] > profiles/xxx.md
"""
import argparse
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maf_yolo_b200 import engine as E  # noqa: E402
from maf_yolo_b200 import topology as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--variant", default="n")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--forward", type=int, default=1, help="which forward pass in the capture (0-based)")
    ap.add_argument("--peak", type=float, default=6542.4)
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
                continue  # multi-metric capture: keep one row per launch
            if d.get("Metric Unit") == "ns":
                d["Metric Value"] = str(float(d["Metric Value"].replace(",", "")))
            elif d.get("Metric Unit") == "us":
                d["Metric Value"] = str(float(d["Metric Value"].replace(",", "")) * 1e3)
            data.append(d)
    names = [re.sub(r"\(.*", "", d["Kernel Name"]) for d in data]
    starts = [i for i, n in enumerate(names) if "stem_conv" in n]
    start = starts[min(a.forward, len(starts) - 1)]
    plan = E.Plan(T.build_graph(a.variant), 640, 640)
    kernel_ops = [o for o in plan.ops if o.kind != "detect_reset"]  # the counter reset is a memset, not a kernel
    ours = [d for d in data[start:] if "mafb200" in d["Kernel Name"]]
    print(f"# per-launch table: MAF-YOLO-{a.variant.upper()} bs={a.batch}, one eager forward "
          f"(ncu gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n")
    print("| op | kernel | grid | us | alg GB/s | frac of HBM peak | TFLOP/s | ideal us @peak |")
    print("|---|---|---|---|---|---|---|---|")
    tot = 0.0
    fam = collections.OrderedDict()
    for op, d in zip(kernel_ops, ours):
        t = float(d["Metric Value"].replace(",", "")) / 1e3
        tot += t
        gb = op.bytes_per_image * a.batch / 1e9
        fl = op.flops_per_image * a.batch / 1e12
        k = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void mafb200::", "").replace("mafb200::", "")
        print(f"| {op.name} | {k} | {d['Grid Size']} | {t:.1f} | {gb / (t / 1e6):.0f} | {gb / (t / 1e6) / a.peak:.3f} | "
              f"{fl / (t / 1e6):.1f} | {gb / a.peak * 1e6:.1f} |")
        f = fam.setdefault(op.kind, [0, 0.0, 0.0, 0.0])
        f[0] += 1
        f[1] += t
        f[2] += gb
        f[3] += fl
    print(f"\ntotal forward: {tot:.1f} us ({a.batch / tot * 1e6:.0f} images/s if launches were back to back)\n")
    print("| family | launches | us | share | alg GB/s | frac of HBM peak | TFLOP/s |")
    print("|---|---|---|---|---|---|---|")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {f[0]} | {f[1]:.1f} | {f[1] / tot:.3f} | {f[2] / (f[1] / 1e6):.0f} | {f[2] / (f[1] / 1e6) / a.peak:.3f} | {f[3] / (f[1] / 1e6):.1f} |")
    rest = [d for d in data[start:] if "mafb200" in d["Kernel Name"]][len(kernel_ops):len(kernel_ops) + 2]
    for d in rest:
        nm = d["Kernel Name"].split("(")[0]
        print(f"\n(next launch: {nm} {float(d['Metric Value'].replace(',', '')) / 1e3:.1f} us)")


if __name__ == "__main__":
    main()
