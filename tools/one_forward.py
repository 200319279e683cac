#!/usr/bin/env python3
"""One eager forward (+NMS) of a variant after a warm-up pass — the command ncu wraps.

  ncu --set full -k regex:dwconv_kernel -s 16 -c 2 ... python tools/one_forward.py --variant n --batch 32
  (`-s` counts only launches matching `-k`; the warm-up forward launches each kernel once.)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import maf_yolo_b200 as mb  # noqa: E402
from maf_yolo_b200 import synth, topology  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="n")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--passes", type=int, default=1)
ap.add_argument("--reference-shaped", action="store_true",
                help="model(x) + non_max_suppression (materialises pred) instead of the serving call bench.py times")
a = ap.parse_args()
g = topology.build_graph(a.variant)
model = mb.from_state_dict(synth.random_state_dict(g, 0), a.variant, use_cuda_graph=False)
x = torch.rand(a.batch, 3, 640, 640, device="cuda")
for _ in range(1 + a.passes):
    if a.reference_shaped:
        pred = model(x)[0]
        mb.non_max_suppression_padded(pred, 0.03, 0.65, multi_label=True, max_det=300)
    else:  # the serving path: decode + candidate filter in the head GEMM epilogues, then one nms_select kernel
        t = model.detect_async(x, 0.03, 0.65, multi_label=True, max_det=300)
        t.done.synchronize()
torch.cuda.synchronize()
print("ok")
