#!/usr/bin/env python3
"""Experiment: where does the e2e gap come from? device-resident fp32 vs uint8 input, with/without H2D."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maf_yolo_b200 as mb
from maf_yolo_b200 import synth, topology
EVAL = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)
g = topology.build_graph("n"); sd = synth.random_state_dict(g, seed=0); dev = torch.device("cuda")
model = mb.from_state_dict(sd, "n", in_flight=2)
host = [torch.randint(0, 256, (32, 3, 640, 640), dtype=torch.uint8).pin_memory() for _ in range(2)]
xu = [h.to(dev) for h in host] + [h.to(dev) for h in host]
xf = [(x.float() / 255) for x in xu[:2]]
def run(name, fn, steps=100):
    for k in range(8): fn(k)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps): fn(k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {32 * steps / dt:.0f} img/s, {1e3 * dt / steps:.3f} ms per step")
run("fp32 resident", lambda k: model.detect_async(xf[k % 2], **EVAL))
run("uint8 resident", lambda k: model.detect_async(xu[k % 4], **EVAL))
cs = torch.cuda.Stream(); ready = [torch.cuda.Event() for _ in range(4)]; cons = [None] * 4
def h2d_step(k):
    j = k % 4
    with torch.cuda.stream(cs):
        if cons[j] is not None: cs.wait_event(cons[j])
        xu[j].copy_(host[k % 2], non_blocking=True); ready[j].record(cs)
    torch.cuda.current_stream().wait_event(ready[j])
    t = model.detect_async(xu[j], **EVAL); cons[j] = t.consumed
run("uint8 + H2D (no D2H)", h2d_step)
# pure H2D bandwidth of one pinned uint8 batch
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(50):
    xu[k % 4].copy_(host[k % 2], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D alone: {host[0].numel() * 50 / dt / 1e9:.1f} GB/s, {1e3 * dt / 50:.3f} ms per 39 MB batch")
# does independent H2D traffic slow the kernels?  (no dependencies between the copies and the forward)
scratch = torch.empty_like(xu[0])
def bg_step(k):
    with torch.cuda.stream(cs):
        scratch.copy_(host[k % 2], non_blocking=True)
    model.detect_async(xu[k % 4], **EVAL)
run("uint8 resident + independent H2D of 39 MB per step", bg_step)
# H2D bandwidth WHILE the forward loop saturates the GPU
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(8): model.detect_async(xu[k % 4], **EVAL)
with torch.cuda.stream(cs):
    e0.record(cs)
for k in range(60):
    with torch.cuda.stream(cs):
        scratch.copy_(host[k % 2], non_blocking=True)
    model.detect_async(xu[k % 4], **EVAL)
with torch.cuda.stream(cs):
    e1.record(cs)
torch.cuda.synchronize()
print(f"H2D under load: {host[0].numel() * 60 / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s ({e0.elapsed_time(e1) / 60:.3f} ms per 39 MB batch on the copy stream)")
