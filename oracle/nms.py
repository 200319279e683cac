"""ORACLE (test infrastructure, not product): CPU restatement of the reference's batched NMS.

Follows yolov6/utils/nms.py:21-105 line by line and restates `torchvision.ops.nms` (torchvision
0.26 CPU kernel `nms_kernel_impl`: stable descending sort, greedy suppression, IoU =
inter / (area_i + area_j - inter) in fp32, threshold compared in double, strict '>').
Pure numpy fp32 — index work must be bit-exact.  Pinned against the reference's own
`non_max_suppression` (imported from /root/reference in this container; tests/test_oracle_cpu.py)
and against the committed golden vectors (tests/golden/).  Parity is otherwise unpinned: the
reference ships no tests or golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np

MAX_WH = 4096  # nms.py:54
MAX_NMS = 30000  # nms.py:55


def xywh2xyxy(x: np.ndarray) -> np.ndarray:
    """nms.py:21-28 (fp32: x - w/2 ...)."""
    y = x.copy()
    two = np.float32(2)
    y[:, 0] = x[:, 0] - x[:, 2] / two
    y[:, 1] = x[:, 1] - x[:, 3] / two
    y[:, 2] = x[:, 0] + x[:, 2] / two
    y[:, 3] = x[:, 1] + x[:, 3] / two
    return y


def nms_indices(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """torchvision.ops.nms on CPU, restated.  boxes [n,4] fp32 xyxy, scores [n] fp32 -> kept indices
    (int64) in descending-score order."""
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    boxes = boxes.astype(np.float32, copy=False)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores.astype(np.float32), kind="stable")  # ties: lower index first
    x1o, y1o, x2o, y2o, ao = x1[order], y1[order], x2[order], y2[order], areas[order]
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float64(iou_thres)
    zero = np.float32(0)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(x1o[i], x1o[i + 1:])
        yy1 = np.maximum(y1o[i], y1o[i + 1:])
        xx2 = np.minimum(x2o[i], x2o[i + 1:])
        yy2 = np.minimum(y2o[i], y2o[i + 1:])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (ao[i] + ao[i + 1:] - inter)
        suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction: np.ndarray, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                        agnostic: bool = False, multi_label: bool = False, max_det: int = 300,
                        max_nms: int = MAX_NMS) -> list[np.ndarray]:
    """nms.py:31-105 without the wall-clock limit.  prediction [B, A, 5+nc] fp32 ->
    list of B arrays [n_i, 6] (x1, y1, x2, y2, score, class) fp32."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    prediction = np.asarray(prediction, dtype=np.float32)
    num_classes = prediction.shape[2] - 5
    conf = np.float32(conf_thres)  # torch compares an fp32 tensor with the scalar cast to fp32
    cand = (prediction[..., 4] > conf) & (prediction[..., 5:].max(axis=-1) > conf)  # nms.py:48
    multi_label = multi_label and num_classes > 1  # nms.py:57
    out = [np.zeros((0, 6), dtype=np.float32) for _ in range(prediction.shape[0])]
    for bi in range(prediction.shape[0]):
        x = prediction[bi][cand[bi]].copy()  # nms.py:62
        if not x.shape[0]:
            continue
        x[:, 5:] *= x[:, 4:5]  # nms.py:69
        box = xywh2xyxy(x[:, :4])  # nms.py:72
        if multi_label:  # nms.py:75-77, rows ordered (anchor asc, class asc)
            bidx, cidx = np.nonzero(x[:, 5:] > conf)
            x = np.concatenate([box[bidx], x[bidx, cidx + 5, None], cidx[:, None].astype(np.float32)], axis=1)
        else:  # nms.py:78-80, first maximum on ties
            cidx = x[:, 5:].argmax(axis=1)
            cf = x[np.arange(x.shape[0]), cidx + 5]
            x = np.concatenate([box, cf[:, None], cidx[:, None].astype(np.float32)], axis=1)[cf > conf]
        if classes is not None:  # nms.py:83-84
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:  # nms.py:90-91 (stable: ties keep the lower row)
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        off = x[:, 5:6] * np.float32(0 if agnostic else MAX_WH)  # nms.py:94
        keep = nms_indices(x[:, :4] + off, x[:, 4], iou_thres)  # nms.py:95-96
        out[bi] = x[keep[:max_det]]  # nms.py:97-100
    return out


def non_max_suppression_tv(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                           agnostic: bool = False, multi_label: bool = False, max_det: int = 300, max_nms: int = MAX_NMS):
    """The same restatement of nms.py:31-105 on torch tensors (CPU or CUDA) with the greedy step done by the library the
    reference itself calls, `torchvision.ops.nms` (nms.py:96) — the form bench.py TIMES as the reference's NMS (the numpy
    form above is ~3x slower than torchvision's C++ kernel and would flatter the GPU/CPU ratio).  Pinned against the
    numpy form in tests/test_oracle_cpu.py.  Returns a list of [n_i, 6] tensors."""
    import torch
    import torchvision

    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    nc = prediction.shape[2] - 5
    cand = (prediction[..., 4] > conf_thres) & (prediction[..., 5:].amax(-1) > conf_thres)  # nms.py:48
    multi_label = multi_label and nc > 1
    out = [torch.zeros((0, 6), device=prediction.device)] * prediction.shape[0]
    for bi in range(prediction.shape[0]):
        x = prediction[bi][cand[bi]]
        if not x.shape[0]:
            continue
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]  # nms.py:69
        half = x[:, 2:4] / 2
        box = torch.cat((x[:, 0:2] - half, x[:, 0:2] + half), 1)  # nms.py:21-28
        if multi_label:
            bidx, cidx = (x[:, 5:] > conf_thres).nonzero(as_tuple=False).T
            x = torch.cat((box[bidx], x[bidx, cidx + 5, None], cidx[:, None].float()), 1)
        else:
            cf, cidx = x[:, 5:].max(1, keepdim=True)
            x = torch.cat((box, cf, cidx.float()), 1)[cf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        if not x.shape[0]:
            continue
        if x.shape[0] > max_nms:
            x = x[x[:, 4].argsort(descending=True, stable=True)[:max_nms]]
        off = x[:, 5:6] * (0 if agnostic else MAX_WH)
        keep = torchvision.ops.nms(x[:, :4] + off, x[:, 4], iou_thres)
        out[bi] = x[keep[:max_det]]
    return out
