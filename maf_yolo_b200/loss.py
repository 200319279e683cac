"""Host side of the training-side kernels: the reference's `ComputeLoss` call for MAF-YOLO's head
(yolov6/models/loss.py:15-162: varifocal + GIoU + DFL on task-aligned assignments) behind the same constructor and
call signature, running as the CUDA kernels of csrc/loss.cu (`mafb200_detect_loss`).

    criterion = maf_yolo_b200.loss.ComputeLoss(num_classes=80, ori_img_size=640)
    loss, loss_items = criterion((feats, pred_scores, pred_distri), targets, epoch_num, step_num)
    loss.backward()          # gradients reach pred_scores / pred_distri (computed by the same kernels)

`outputs` is what `Detect_yaml.forward` returns in train mode (yolov6/models/yolo.py:333-354): `feats` is only used for
its device in the reference and is ignored here; `pred_scores` [B,A,nc] are class probabilities, `pred_distri`
[B,A,68] DFL logits.  `targets` is the collated [T,6] tensor (image, class, cx, cy, w, h).  Like the reference the
loss is float64 (its target tensor is float64, loss.py:165-169) and `loss_items` = (2.5 iou, 0.5 dfl, 1.0 cls), detached.

Both assigners of the reference are implemented: ATSS (yolov6/assigners/atss_assigner.py) while `epoch_num < warmup_epoch`,
the task-aligned assigner afterwards (loss.py:83-100).  One deviation: in the warm-up epochs the reference's class loss is
an fp32 sum (its target scores become fp32 there); this path accumulates it in float64 like the other terms (relative
difference ~1e-7).  There is no CPU path: tensors must live on an sm_100 GPU.
"""
from __future__ import annotations

import torch

from . import _lib


class _DetectLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_scores, pred_distri, targets, img_size, num_classes, gt_cap, owner, assigner):
        ps = pred_scores.detach().float().contiguous()
        pd = pred_distri.detach().float().contiguous()
        b, a, nc = ps.shape
        dev = ps.device
        tg = targets.detach().to(device=dev, dtype=torch.float32).contiguous().view(-1, 6)
        ws_bytes = _lib.lib().mafb200_loss_workspace_bytes(b, a, gt_cap)
        ws = owner._workspace(dev, ws_bytes)
        scalars = torch.empty(8, dtype=torch.float64, device=dev)
        need_grad = pred_scores.requires_grad or pred_distri.requires_grad
        gs = torch.empty_like(ps) if need_grad else None
        gd = torch.empty_like(pd) if need_grad else None
        gt_idx = torch.empty(b, a, dtype=torch.int32, device=dev)
        fg = torch.empty(b, a, dtype=torch.uint8, device=dev)
        tscore = torch.empty(b, a, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().mafb200_detect_loss(
            ps.data_ptr(), pd.data_ptr(), tg.data_ptr() if tg.numel() else None, tg.shape[0], b, img_size, num_classes, gt_cap,
            owner._boxes_override.data_ptr() if owner._boxes_override is not None else None, ws.data_ptr(), ws_bytes,
            scalars.data_ptr(), gs.data_ptr() if need_grad else None, gd.data_ptr() if need_grad else None,
            gt_idx.data_ptr(), fg.data_ptr(), tscore.data_ptr(), assigner, torch.cuda.current_stream(dev).cuda_stream))
        owner.last = dict(scalars=scalars, target_gt_idx=gt_idx, fg_mask=fg, target_score=tscore)
        ctx.save_for_backward(gs, gd)
        ctx.in_dtypes = (pred_scores.dtype, pred_distri.dtype)
        ctx.mark_non_differentiable(scalars)
        return scalars[0].clone(), scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        gs, gd = ctx.saved_tensors
        g = g_loss.to(torch.float32)
        return ((gs * g).to(ctx.in_dtypes[0]) if gs is not None else None,
                (gd * g).to(ctx.in_dtypes[1]) if gd is not None else None, None, None, None, None, None, None)


class ComputeLoss:
    """Drop-in for `yolov6.models.loss.ComputeLoss` (same constructor arguments and call)."""

    def __init__(self, fpn_strides=(8, 16, 32), grid_cell_size=5.0, grid_cell_offset=0.5, num_classes=80, ori_img_size=640,
                 warmup_epoch=3, use_dfl=True, reg_max=16, iou_type="giou", loss_weight=None):
        if tuple(fpn_strides) != (8, 16, 32) or grid_cell_offset != 0.5 or not use_dfl or reg_max != 16 or iou_type != "giou":
            raise ValueError("mafb200 ComputeLoss implements MAF-YOLO's configuration: strides (8,16,32), offset 0.5, "
                             "DFL with reg_max 16, GIoU (configs/*.py of the reference)")
        lw = loss_weight or {"class": 1.0, "iou": 2.5, "dfl": 0.5}
        if (lw["class"], lw["iou"], lw["dfl"]) != (1.0, 2.5, 0.5):
            raise ValueError("loss weights are fixed to the reference's {'class': 1.0, 'iou': 2.5, 'dfl': 0.5} (loss.py:32-35)")
        self.num_classes = int(num_classes)
        self.ori_img_size = int(ori_img_size)
        self.warmup_epoch = int(warmup_epoch)
        self.loss_weight = lw
        self._ws = {}
        self._boxes_override = None  # tests: boxes [B,A,4] (stride units) instead of decoding pred_distri
        self.last = None             # assignment of the most recent call (device tensors)

    def _workspace(self, dev, nbytes):
        ws = self._ws.get(dev)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ws[dev] = ws
        return ws

    def __call__(self, outputs, targets, epoch_num, step_num, gt_cap=None):
        _feats, pred_scores, pred_distri = outputs
        assigner = 1 if epoch_num < self.warmup_epoch else 0  # loss.py:83: ATSS while warming up, task-aligned afterwards
        if not pred_scores.is_cuda:
            raise RuntimeError("mafb200 ComputeLoss has no CPU path: predictions must be CUDA tensors")
        if pred_scores.type() != pred_distri.type():
            raise AssertionError("pred_scores and pred_distri must have the same type")  # loss.py:68
        if gt_cap is None:  # largest number of boxes in one image (one small device -> host read, as the reference's .cpu())
            t = targets.view(-1, 6)
            gt_cap = int(torch.bincount(t[:, 0].long(), minlength=1).max().item()) if t.shape[0] else 0
        loss, scalars = _DetectLossFn.apply(pred_scores, pred_distri, targets, self.ori_img_size, self.num_classes, int(gt_cap), self,
                                            assigner)
        return loss, scalars[1:4].detach()
