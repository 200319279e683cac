"""Reference-facing host API: drop-ins with the reference's own names, arguments and error behaviour.

  non_max_suppression(...)    == yolov6/utils/nms.py:31 (same signature, same results, list of [n,6])
  convert(model) / B200DetectModel.forward(x) -> [pred, featmaps]   == yolov6/models/yolo.py:179-209
  block drop-ins (RepVGGBlock, ConvWrapper, RepHDW, MPRep, SPPF, Head_DepthUni, Detect_yaml): engine.py

Everything computes through libmafb200.so; PyTorch only owns memory and streams.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops

_MAX_NMS = 30000  # yolov6/utils/nms.py:55
_ws_cache: dict = {}


def _nms_buffers(device, b: int, a: int, nc: int, max_det: int):
    key = (str(device), b, a, nc, max_det)
    hit = _ws_cache.get(key)
    if hit is None:
        nbytes = ops.nms_workspace_bytes(b, a, nc)
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device)
        hit = (ws,)
        _ws_cache.clear()  # keep one shape resident (the workspace can be hundreds of MB)
        _ws_cache[key] = hit
    return hit[0]


def non_max_suppression_padded(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                               classes: Optional[Sequence[int]] = None, agnostic: bool = False,
                               multi_label: bool = False, max_det: int = 300, max_nms: int = _MAX_NMS,
                               det: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None):
    """Sync-free form: returns (det [B,max_det,6] fp32, count [B] int32), both on the device.
    Fixed-size outputs make it CUDA-graph capturable and all-gatherable (maf_yolo_b200.dist)."""
    # same checks, same messages as nms.py:50-51
    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'
    if not prediction.is_cuda:
        raise RuntimeError("maf_yolo_b200.non_max_suppression needs a CUDA tensor (no CPU fallback)")
    pred = prediction
    if pred.dtype != torch.float32 or not pred.is_contiguous():
        pred = pred.float().contiguous()
    b, a, no = pred.shape
    nc = no - 5
    with torch.cuda.device(pred.device):
        ws = _nms_buffers(pred.device, b, a, nc, max_det)
        if det is None:
            det = torch.empty((b, max_det, 6), dtype=torch.float32, device=pred.device)
        if count is None:
            count = torch.empty((b,), dtype=torch.int32, device=pred.device)
        filt = None
        if classes is not None:
            filt = torch.zeros(nc, dtype=torch.uint8)
            for c in classes:
                if 0 <= int(c) < nc:
                    filt[int(c)] = 1
            filt = filt.to(pred.device)
        ops.nms(pred, conf_thres, iou_thres, multi_label, agnostic, filt, max_det, max_nms, det, count, ws)
    return det, count


def non_max_suppression(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                        classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                        max_det: int = 300, max_nms: int = _MAX_NMS) -> List[torch.Tensor]:
    """Drop-in for yolov6/utils/nms.py:31 — list of per-image [n, 6] (xyxy, conf, cls) device tensors."""
    det, count = non_max_suppression_padded(prediction, conf_thres, iou_thres, classes, agnostic, multi_label,
                                            max_det, max_nms)
    counts = count.tolist()  # the one device->host sync (the reference syncs once per image)
    return [det[i, :n] for i, n in enumerate(counts)]
