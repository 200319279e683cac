"""Reference-facing host API: drop-ins with the reference's own names, arguments and error behaviour.

  non_max_suppression(...)    == yolov6/utils/nms.py:31 (same signature, same results, list of [n,6])
  convert(model) / B200DetectModel.forward(x) -> [pred, featmaps]   == yolov6/models/yolo.py:179-209
  block drop-ins (RepVGGBlock, ConvWrapper, RepHDW, MPRep, SPPF, Head_DepthUni, Detect_yaml): engine.py

Everything computes through libmafb200.so; PyTorch only owns memory and streams.
"""
from __future__ import annotations

import os
from typing import List, NamedTuple, Optional, Sequence

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
                               det: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None,
                               workspace: Optional[torch.Tensor] = None):
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
        # the shared cached workspace is only safe for calls ordered on one stream; concurrent callers
        # (detect_async with several batches in flight) pass their own
        ws = workspace if workspace is not None else _nms_buffers(pred.device, b, a, nc, max_det)
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


# ------------------------------------------------------------------------------------------------
# whole-model drop-in
# ------------------------------------------------------------------------------------------------
class B200DetectModel(torch.nn.Module):
    """Drop-in for the reference `Model` (yolov6/models/yolo.py:122-217) in eval mode.

    forward(x, val_loss=False) -> [pred, featmaps]:  pred is the `[B, A, 5+nc]` fp32 tensor of
    Detect_yaml's eval branch (yolo.py:355-396); featmaps — unused by every caller
    (yolov6/core/evaler.py:168, yolov6/layers/common.py:368) — is [] unless `return_featmaps`.
    forward(x, val_loss=True) -> [(feats, pred_scores [B,A,nc], pred_distri [B,A,68]), []]: the train-form outputs of
    Detect_yaml (yolo.py:333-354) with frozen BN — the input of `maf_yolo_b200.loss.ComputeLoss` (validation loss).
    `.train()` raises: batch-stat BN and the backward pass are not built.
    Keeps the attributes callers read: `stride`, `nc`, `names` (evaler.py:94,151,246).
    `.half()` / `.float()` (evaler.py:112) are accepted and change nothing: the compute type is
    fixed (fp16 operands, fp32 accumulate) and the output is fp32, as the reference's is.
    Engines (plan + arena + CUDA graph) are built lazily per (batch, H, W, device).
    Output lifetime: `pred` is a fresh tensor per call (as in the reference) unless the model was built with
    `borrow_output=True`, in which case it is one of two engine-owned buffers and stays valid until the second next
    forward() of the same input shape.  `detect_async` results have their own documented lifetime (DetectTicket).
    """

    def __init__(self, graph, folded, names=None, use_cuda_graph: bool = True, return_featmaps: bool = False,
                 n_streams: int = 4, in_flight: int = 1, borrow_output: bool = False):
        super().__init__()
        self.graph = graph
        self.folded = folded
        self.nc = graph.nc
        self.names = names if names is not None else [str(i) for i in range(graph.nc)]
        self.stride = torch.tensor(graph.strides)
        self.use_cuda_graph = use_cuda_graph
        self.return_featmaps = return_featmaps
        self.n_streams = n_streams
        # forward() returns a fresh tensor like the reference's Model.forward does (a caller may keep outputs across
        # batches: `outs.append(model(x)[0])`).  borrow_output=True hands out the engine-owned buffer instead — no
        # 2.9 MB/image copy — which is overwritten by the SECOND next forward() of the same (batch, H, W).
        self.borrow_output = borrow_output
        # detect_async: decode fused with the NMS threshold / compaction pass (MAFB200_FUSED_DETECT=0 turns it off)
        self.fused_detect = os.environ.get("MAFB200_FUSED_DETECT", "1") != "0"
        self.in_flight = max(1, int(in_flight))  # detect_async: engine replicas (arena + graph + streams) used round-robin
        self._rr = 0
        self._engines = {}
        self.training = False

    # the reference evaler calls these; weights are already packed, nothing to cast
    def half(self):
        return self

    def float(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("B200DetectModel is inference-only (deploy-form folded weights)")
        return super().train(False)

    def engine_for(self, x: torch.Tensor, slot: int = 0):
        from .engine import Engine

        b, _, h, w = x.shape
        key = (b, h, w, str(x.device), slot)
        eng = self._engines.get(key)
        if eng is None:
            # return_featmaps needs the heads' cls / reg maps in memory: the round-1 form without the K7 epilogues
            eng = Engine(self.graph, self.folded, b, h, w, x.device, self.use_cuda_graph, n_streams=self.n_streams,
                         k7=False if self.return_featmaps else None)
            self._engines[key] = eng
        return eng

    @torch.no_grad()
    def forward(self, x: torch.Tensor, val_loss: bool = False):
        if not x.is_cuda:
            raise RuntimeError("B200DetectModel needs a CUDA input tensor: there is no CPU fallback on this path")
        eng = self.engine_for(x)
        if val_loss:
            # Detect_yaml.forward's `self.training or val_loss` branch (yolo.py:333-354) in eval mode (BN frozen = the folded
            # weights): (feats, pred_scores [B,A,nc] probabilities, pred_distri [B,A,68] DFL logits) — what ComputeLoss takes
            # (the trainer's validation loss).  The two tensors come from the cls_pred / reg_pred GEMM epilogues in fp32.
            with torch.cuda.device(x.device):
                scores, distri = eng.forward(x, detect="train")
                scores, distri = scores.clone(), distri.clone()
                if eng.reuse_buffers:  # the level stems' arena ranges may have been handed to later layers
                    feats = [torch.zeros((x.shape[0], lv[0].c, lv[0].buf.h, lv[0].buf.w), device=x.device) for lv in eng.plan.level_views]
                else:
                    feats = [eng.view(lv[0]).to_nchw() for lv in eng.plan.level_views]  # Head_DepthUni's `x` = stem output
            return [(feats, scores, distri), []]
        with torch.cuda.device(x.device):
            pred = eng.forward(x)
            if not self.borrow_output:
                pred = pred.clone()
            feats = eng.head_outputs() if self.return_featmaps else []
        return [pred, feats]


    @torch.no_grad()
    def detect_async(self, x: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                     classes: Optional[Sequence[int]] = None, agnostic: bool = False, multi_label: bool = False,
                     max_det: int = 300, max_nms: int = _MAX_NMS, after_nms=None, gather=None) -> "DetectTicket":
        """Serving form of `pred = model(x)[0]; non_max_suppression(pred, ...)` (evaler.py:168,178) that keeps
        several batches in flight instead of draining the GPU between them:
          * the NMS of a batch runs on a side stream and overlaps the forward of the next call (its 32 CTAs
            leave most SMs idle), predictions are double-buffered;
          * with `in_flight` = 2 (constructor) consecutive calls alternate between two engine replicas (own
            arena, CUDA graph and streams), so the latency-bound low-resolution tail of one forward overlaps
            the bandwidth-bound head of the next (measured on B200: 14.5k -> 15.2k -> 17.1k images/s).
        Nothing is synchronised with the host.  Returns a DetectTicket: wait on `.done` (event) before reading
        `.det` [B,max_det,6] / `.count` [B] — the buffers are reused 2 * in_flight calls later; `.consumed` fires
        when `x` has been read.  `after_nms(det, count)` — optional — is invoked on the NMS stream right after
        the NMS (e.g. to enqueue a D2H copy); its result is `.extra`.
        `gather` — a maf_yolo_b200.dist.DetectionGather with 2 * in_flight copies — makes the NMS write this rank's rows
        of the multi-GPU gather buffer in place and enqueues the ONE all-gather behind it; `.det` / `.count` are then
        views of the whole job's batch ([world * B, max_det, 6], [world * B])."""
        if not x.is_cuda:
            raise RuntimeError("B200DetectModel needs a CUDA input tensor: there is no CPU fallback on this path")
        slot = self._rr % self.in_flight
        self._rr += 1
        eng = self.engine_for(x, slot)
        dev = x.device
        with torch.cuda.device(dev):
            st = getattr(eng, "_async_state", None)
            if st is None or st["max_det"] != max_det:
                b, a, nc = eng.batch, eng.plan.anchors, self.nc
                st = {"max_det": max_det, "stream": torch.cuda.Stream(device=dev),
                      "fwd": torch.cuda.Stream(device=dev) if self.in_flight > 1 else None,
                      "det": [torch.empty((b, max_det, 6), dtype=torch.float32, device=dev) for _ in range(2)],
                      "cnt": [torch.empty((b,), dtype=torch.int32, device=dev) for _ in range(2)],
                      "done": [torch.cuda.Event() for _ in range(2)],
                      "ws": None if self.fused_detect else
                      torch.empty((ops.nms_workspace_bytes(b, a, nc) + 7) // 8, dtype=torch.int64, device=dev)}
                eng._async_state = st
            caller = torch.cuda.current_stream(dev)
            fwd_stream = st["fwd"] if st["fwd"] is not None else caller
            if fwd_stream is not caller:
                ready = torch.cuda.Event()
                ready.record(caller)
                fwd_stream.wait_event(ready)  # x was produced on the caller's stream
                x.record_stream(fwd_stream)
            with torch.cuda.stream(fwd_stream):
                k = eng._flip  # the prediction buffer this call will write (forward waits for its last reader)
                fused = self.fused_detect and max_nms > 0
                if fused:
                    # decode + NMS threshold/compaction in one kernel; the [B, A, 5+nc] tensor is never written
                    boxes = eng.forward(x, detect=(float(conf_thres), bool(multi_label),
                                                   tuple(int(c) for c in classes) if classes is not None else None))
                else:
                    pred = eng.forward(x)
                fwd_done = torch.cuda.Event()
                fwd_done.record(fwd_stream)
                eng.last_async_forward = fwd_done
            side = st["stream"]
            side.wait_event(fwd_done)
            with torch.cuda.stream(side):
                gi = 2 * slot + k  # buffer of the gatherer used by this call
                if gather is not None and (gather.max_det != max_det or gather.batch != eng.batch or len(gather.bufs) < 2 * self.in_flight):
                    raise ValueError("gather: DetectionGather(batch, max_det, copies=2 * in_flight) does not match this call")
                if fused:
                    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
                    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'
                    if gather is not None:
                        ops.nms_select_packed(boxes, self.nc, iou_thres, agnostic, max_det, max_nms, gather.mine(gi), eng.nms_ws[k])
                    else:
                        det, cnt = st["det"][k], st["cnt"][k]
                        ops.nms_select(boxes, self.nc, iou_thres, agnostic, max_det, max_nms, det, cnt, eng.nms_ws[k])
                else:
                    det, cnt = non_max_suppression_padded(pred, conf_thres, iou_thres, classes, agnostic, multi_label,
                                                          max_det, max_nms, det=st["det"][k], count=st["cnt"][k],
                                                          workspace=st["ws"])
                    if gather is not None:  # unfused path: stage through the private buffers, then into the gather rows
                        mine = gather.mine(gi)
                        mine[:, :max_det * 6].copy_(det.reshape(det.shape[0], -1))
                        mine.view(torch.int32)[:, max_det * 6].copy_(cnt)
                if gather is not None:
                    gather.gather(gi)
                    det, cnt = gather.views(gi)
                    # with a gatherer the callback also gets this rank's packed rows (contiguous [B, max_det*6+2]:
                    # detections + count bits), so that a D2H of the local results is ONE copy
                    extra = after_nms(det, cnt, gather.mine(gi)) if after_nms is not None else None
                else:
                    extra = after_nms(det, cnt) if after_nms is not None else None
                st["done"][k].record(side)
            eng.reader_done[k] = st["done"][k]
        return DetectTicket(det, cnt, st["done"][k], fwd_done, extra)


class DetectTicket(NamedTuple):
    """Result handle of B200DetectModel.detect_async (all device-side; see its docstring)."""
    det: torch.Tensor
    count: torch.Tensor
    done: "torch.cuda.Event"
    consumed: "torch.cuda.Event"
    extra: object = None


def from_state_dict(state_dict, variant_or_yaml="n", nc: int = 80, bn_eps: float = 1e-3, **kw) -> B200DetectModel:
    """Builds the drop-in from reference weights (train- or deploy-form `state_dict`)."""
    from .fold import fold_state_dict
    from .topology import build_graph

    graph = build_graph(variant_or_yaml, nc)
    return B200DetectModel(graph, fold_state_dict(graph, state_dict, bn_eps), **kw)


def convert(model: torch.nn.Module, **kw) -> B200DetectModel:
    """Post-construction model surgery, the reference's own extension idiom (evaler.py:101-109):
    takes a live reference `Model` (train or deploy form, any device) and returns the B200 drop-in.
    Reads the topology from `model.yaml` and BatchNorm eps from the modules."""
    rows = getattr(model, "yaml", None)
    if rows is None:
        raise TypeError("convert() expects a yaml-built reference Model (model.yaml missing)")
    nc = getattr(getattr(model, "detect", None), "nc", None) or kw.pop("nc", 80)
    eps = 1e-3
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            eps = m.eps
            break
    names = getattr(model, "names", None)
    return from_state_dict(model.state_dict(), rows, nc, eps, names=names, **kw)
