#!/usr/bin/env python3
"""Experiment: two full-batch engines on two streams, alternate batches (two forwards in flight)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maf_yolo_b200 as mb
from maf_yolo_b200 import synth, topology
EVAL = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
g = topology.build_graph("n"); sd = synth.random_state_dict(g, seed=0); dev = torch.device("cuda")
x32 = [torch.rand(32, 3, 640, 640, device=dev) for _ in range(2)]
for depth in (2, 3, 4):
    nst = 1
    models = [mb.from_state_dict(sd, "n", n_streams=nst) for _ in range(depth)]
    streams = [torch.cuda.Stream() for _ in range(depth)]
    def step(k):
        i = k % depth
        with torch.cuda.stream(streams[i]):
            models[i].detect_async(x32[i % 2], **EVAL)
    for k in range(8): step(k)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(80): step(k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{depth} engines alternating, graph_streams={nst}: {32 * 80 / dt:.0f} img/s, {1e3 * dt / 80:.3f} ms per step")
    del models
