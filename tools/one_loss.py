#!/usr/bin/env python3
"""One call of the training-side kernels (mafb200_detect_loss through maf_yolo_b200.loss.ComputeLoss) after a warm-up
call — the command ncu wraps for the loss path's launch list (batch 16 = the reference's per-GPU training batch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from maf_yolo_b200.loss import ComputeLoss  # noqa: E402

batch, boxes = int(os.environ.get("BS", "16")), 20
g = torch.Generator().manual_seed(3)
scores = torch.sigmoid(torch.randn(batch, 8400, 80, generator=g) * 1.5 - 3.0).cuda().requires_grad_()
distri = torch.randn(batch, 8400, 68, generator=g).cuda().requires_grad_()
rows = []
for b in range(batch):
    for _ in range(boxes):
        cx, cy = torch.rand(2, generator=g).tolist()
        w, h = (0.04 + 0.45 * torch.rand(2, generator=g)).tolist()
        rows.append([float(b), float(int(torch.randint(0, 80, (1,), generator=g))), cx, cy, w, h])
targets = torch.tensor(rows).cuda()
crit = ComputeLoss(warmup_epoch=0)
for _ in range(2):
    loss, items = crit((None, scores, distri), targets, 0, 0, gt_cap=boxes)
torch.cuda.synchronize()
print("ok", loss.item())
