"""Post-NMS step of the reference on the device, behind the reference's own method names.

  convert_to_coco_format(det, count, img_shape, paths, shapes, ids)   == Evaler.convert_to_coco_format
                                                                          (yolov6/core/evaler.py:420-442)
  scale_coords(...)                                                   == Evaler.scale_coords (evaler.py:391-418)
  rescale(ori_shape, det, count, target_shape)                        == Inferer.rescale (yolov6/core/inferer.py:181-195)

The reference loops over images and detections in Python with a `.tolist()` / `.item()` device sync per
detection; here ONE kernel (mafb200_scale_detections) rescales, clips and converts the whole padded batch,
one D2H copy brings it to the host, and only the json-dict packing (Python `round`, exactly the reference's
expression on the same float values) stays on the CPU.

`recip_mul` — which division the reference's `coords /= gain` (evaler.py:404-411) is:
  False (default): IEEE division `a / b`, what torch computes for CPU tensors — bit-exact against the CPU oracle
                   (oracle/postprocess.py) and the committed golden vectors, which is what the parity tests pin;
  True:            `a * (1 / b)`, what torch's CUDA kernel computes when the divisor is a Python scalar, i.e. what the
                   reference evaler produces when it runs on a GPU.  The two differ by at most 1 ulp per coordinate;
                   pass True to reproduce a GPU run of the reference bit for bit.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def scale_detections(det: torch.Tensor, count: torch.Tensor, params: torch.Tensor, mode: str = "xyxy",
                     category_ids: Optional[torch.Tensor] = None, recip_mul: bool = False,
                     out: Optional[torch.Tensor] = None, out_cat: Optional[torch.Tensor] = None):
    """det [B,max_det,6] fp32, count [B] int32, params [B,6] fp32 (gain_x, gain_y, pad_x, pad_y, w0, h0), all CUDA.
    mode 'xyxy' | 'coco' (top-left xywh).  Returns (out, out_cat)."""
    if not det.is_cuda:
        raise RuntimeError("maf_yolo_b200.postprocess needs CUDA tensors (no CPU fallback)")
    assert det.dtype == torch.float32 and det.dim() == 3 and det.shape[2] == 6 and det.is_contiguous()
    assert count.dtype == torch.int32 and params.dtype == torch.float32 and params.is_contiguous()
    b, max_det, _ = det.shape
    assert tuple(params.shape) == (b, 6) and count.numel() == b
    with torch.cuda.device(det.device):
        if out is None:
            out = torch.empty_like(det)
        if out_cat is None:
            out_cat = torch.empty((b, max_det), dtype=torch.int32, device=det.device)
        nc = int(category_ids.numel()) if category_ids is not None else 0
        check(lib().mafb200_scale_detections(det.data_ptr(), count.data_ptr(), b, max_det, params.data_ptr(),
                                             category_ids.data_ptr() if category_ids is not None else None, nc,
                                             {"xyxy": 0, "coco": 1}[mode], int(bool(recip_mul)), out.data_ptr(),
                                             out_cat.data_ptr(), _stream()))
    return out, out_cat


def _params_from_shapes(shapes, scale_exact: bool) -> torch.Tensor:
    """shapes[i] = ((h0, w0), ((ratio_h, ratio_w), (pad_w, pad_h))) as the reference's dataloader yields them
    (evaler.py:424-426).  Python floats become fp32 exactly where torch would cast them."""
    rows = []
    for shape, (gain, pad) in ((s[0], s[1]) for s in shapes):
        gx = gain[1] if scale_exact else gain[0]
        rows.append([gx, gain[0], pad[0], pad[1], shape[1], shape[0]])
    return torch.tensor(rows, dtype=torch.float64).to(torch.float32)


def scale_coords(img1_shape, det: torch.Tensor, count: torch.Tensor, shapes, scale_exact: bool = False,
                 recip_mul: bool = False) -> torch.Tensor:
    """Batched Evaler.scale_coords: returns det with columns 0-3 rescaled to each original image (xyxy)."""
    params = _params_from_shapes(shapes, scale_exact).to(det.device)
    return scale_detections(det, count, params, "xyxy", recip_mul=recip_mul)[0]


def rescale(ori_shape, det: torch.Tensor, count: torch.Tensor, target_shapes: Sequence[Sequence[int]],
            recip_mul: bool = False) -> torch.Tensor:
    """Batched Inferer.rescale (inferer.py:181-195): `ori_shape` = network input (H, W), target_shapes[i] = the
    original image (h0, w0)."""
    rows = []
    for tgt in target_shapes:
        ratio = min(ori_shape[0] / tgt[0], ori_shape[1] / tgt[1])
        rows.append([ratio, ratio, (ori_shape[1] - tgt[1] * ratio) / 2, (ori_shape[0] - tgt[0] * ratio) / 2, tgt[1], tgt[0]])
    params = torch.tensor(rows, dtype=torch.float64).to(torch.float32).to(det.device)
    return scale_detections(det, count, params, "xyxy", recip_mul=recip_mul)[0]


def convert_to_coco_format(det: torch.Tensor, count: torch.Tensor, img_shape, paths, shapes, ids,
                           is_coco: bool = True, scale_exact: bool = False, recip_mul: bool = False):
    """Drop-in for Evaler.convert_to_coco_format (evaler.py:420-442) on the padded NMS output
    (maf_yolo_b200.non_max_suppression_padded): same list of {"image_id","category_id","bbox","score"} dicts."""
    params = _params_from_shapes(shapes, scale_exact).to(det.device)
    ids_t = torch.tensor(list(ids), dtype=torch.int32, device=det.device)
    out, cat = scale_detections(det, count, params, "coco", ids_t, recip_mul)
    out_h, cat_h, cnt_h = out.cpu().numpy(), cat.cpu().numpy(), count.cpu().numpy()  # one sync for the batch
    results = []
    for i, n in enumerate(cnt_h.tolist()):
        if n == 0:
            continue
        path = Path(paths[i])
        image_id = int(path.stem) if is_coco else path.stem
        boxes = out_h[i, :n, :4].astype(np.float64).tolist()  # fp32 -> python float, as tensor.tolist() does
        scores = out_h[i, :n, 4].astype(np.float64).tolist()
        cats = cat_h[i, :n].tolist()
        for bb, sc, c in zip(boxes, scores, cats):
            results.append({"image_id": image_id, "category_id": c, "bbox": [round(x, 3) for x in bb],
                            "score": round(sc, 5)})
    return results
