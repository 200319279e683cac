#!/usr/bin/env python3
"""Experiment: one batch-32 engine vs two concurrent batch-16 engines on two streams (same GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maf_yolo_b200 as mb
from maf_yolo_b200 import synth, topology

EVAL = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
g = topology.build_graph("n")
sd = synth.random_state_dict(g, seed=0)
dev = torch.device("cuda")
x32 = [torch.rand(32, 3, 640, 640, device=dev) for _ in range(2)]

def bench(nsplit, steps=60):
    B = 32 // nsplit
    models = [mb.from_state_dict(sd, "n", use_cuda_graph=True, n_streams=4) for _ in range(nsplit)]
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    dets = [torch.empty((B, 300, 6), device=dev) for _ in range(nsplit)]
    cnts = [torch.empty((B,), dtype=torch.int32, device=dev) for _ in range(nsplit)]
    xs = [[x[i * B:(i + 1) * B].contiguous() for i in range(nsplit)] for x in x32]
    def step(k):
        for i in range(nsplit):
            with torch.cuda.stream(streams[i]):
                pred = models[i](xs[k % 2][i])[0]
                mb.non_max_suppression_padded(pred, **EVAL, det=dets[i], count=cnts[i])
    for k in range(6):
        step(k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        step(k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"nsplit={nsplit}: {32 * steps / dt:.0f} img/s, {1e3 * dt / steps:.3f} ms per 32 images")

for n in (1, 2, 4):
    bench(n)

# ---- experiment 2: NMS of step i on a side stream overlapping forward of step i+1 (hazard ignored: timing only)
model = mb.from_state_dict(sd, "n", use_cuda_graph=True, n_streams=4)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
det = torch.empty((32, 300, 6), device=dev); cnt = torch.empty((32,), dtype=torch.int32, device=dev)
def step2(k):
    with torch.cuda.stream(sa):
        pred = model(x32[k % 2])[0]
        ev = torch.cuda.Event(); ev.record(sa)
    with torch.cuda.stream(sb):
        sb.wait_event(ev)
        mb.non_max_suppression_padded(pred, **EVAL, det=det, count=cnt)
for k in range(6): step2(k)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(60): step2(k)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pipelined NMS: {32 * 60 / dt:.0f} img/s, {1e3 * dt / 60:.3f} ms per step")
