"""Block-level drop-ins: one yaml layer of the reference model as an `nn.Module` with the reference
block's forward signature (NCHW float tensors in and out), computed by the C-ABI kernels.

    RepVGGBlock.forward(inputs) -> T          yolov6/layers/common.py:214-217
    ConvWrapper.forward(x) -> T               common.py:82
    RepHDW.forward(x) -> T                    common.py:938-946
    MPRep.forward(input) -> T                 common.py:787-792
    SPPF.forward(x) -> T                      common.py:123-129
    Head_DepthUni.forward(x) -> (T, T, T)     common.py:1325-1336   (stem, sigmoid(cls), reg)
    Detect_yaml.forward(list, val_loss) -> T  yolov6/models/yolo.py:355-396 (eval branch)

`convert_blocks(model)` performs the reference's own kind of model surgery (evaler.py:101-109): it walks
`model.backbone`, swaps every supported block for a `B200Block` that keeps `.i/.f/.type/.np`
(yolo.py:113,190-200), and swaps `model.detect`, so the unmodified `Model.forward` loop keeps working.
These exist for interchangeability and per-block parity tests; they pay a layout conversion
(NCHW fp32 <-> NHWC fp16) at every block boundary — the fast path is the whole-graph `convert()`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .engine import Engine, Plan
from .fold import Folded, fold_state_dict
from .ops import NHWC
from .topology import Graph, Layer, build_graph

SUPPORTED = {"repvgg": "RepVGGBlock", "convw": "ConvWrapper", "rephdw": "RepHDW", "mprep": "MPRep", "sppf": "SPPF",
             "head": "Head_DepthUni"}


class _BlockEngine(Engine):
    """Engine over a single-layer plan: inputs are converted from NCHW, outputs back to NCHW fp32."""

    def __init__(self, graph: Graph, folded: Folded, layer: int, batch: int, height: int, width: int, device):
        self.device = torch.device(device)
        self.graph, self.batch, self.height, self.width = graph, batch, height, width
        self.plan = Plan(graph, height, width, only_layer=layer, head_sigmoid=True)
        self.use_cuda_graph, self.n_streams = False, 1
        with torch.cuda.device(self.device):
            nbytes = self.plan.assign_offsets(batch, reuse=False)
            self.arena = torch.zeros(max(nbytes // 2, 8), dtype=torch.float16, device=self.device)
            self.pred = None
            self._views, self._weights, self._x = {}, {}, None
            self._calls = [self._bind(op, folded) for op in self.plan.ops]
        self._graph = None
        self._schedule = None
        self._side_streams = []
        self.launches_per_forward = len(self._calls)

    def run(self, inputs: Sequence[torch.Tensor]):
        lay = self.graph.layers[self.plan.only_layer]
        if lay.frm[0] < 0:  # network stem reads the NCHW image directly
            self._x = inputs[0].contiguous()
        else:
            for v, x in zip(self.plan.inputs, inputs):
                ops.nchw_to_nhwc(x.contiguous(), self.view(v))
        for call in self._calls:
            call()
        out = self.plan.layer_out[lay.i]
        outs = out if isinstance(out, tuple) else (out,)
        res = []
        for v in outs:
            nv = self.view(v)
            t = torch.empty((nv.n, nv.c, nv.h, nv.w), dtype=torch.float32, device=self.device)
            ops.nhwc_to_nchw(nv, t)
            res.append(t)
        return tuple(res) if isinstance(out, tuple) else res[0]


class B200Block(torch.nn.Module):
    def __init__(self, graph: Graph, folded: Folded, layer: int):
        super().__init__()
        self.graph, self.folded, self.layer = graph, folded, layer
        lay = graph.layers[layer]
        self.kind = lay.kind
        self.i, self.f = lay.i, (lay.frm[0] if len(lay.frm) == 1 and lay.frm[0] == lay.i - 1 else lay.frm)
        self.type, self.np = f"maf_yolo_b200.{SUPPORTED[lay.kind]}", 0
        self._engines: Dict[Tuple, _BlockEngine] = {}

    @torch.no_grad()
    def forward(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        if not xs[0].is_cuda:
            raise RuntimeError("B200Block needs CUDA tensors: there is no CPU fallback on this path")
        key = (tuple(tuple(t.shape) for t in xs), str(xs[0].device))
        eng = self._engines.get(key)
        if eng is None:
            b, _, h, w = xs[0].shape
            eng = _BlockEngine(self.graph, self.folded, self.layer, b, h, w, xs[0].device)
            self._engines[key] = eng
        with torch.cuda.device(xs[0].device):
            xs = [t if t.dtype in (torch.float32, torch.float16) or self.graph.layers[self.layer].frm[0] < 0 else t.float()
                  for t in xs]
            return eng.run(xs)


class B200Detect(torch.nn.Module):
    """Drop-in for Detect_yaml (eval branch): list of (stem, cls_prob, reg) NCHW -> pred [B, A, 5+nc]."""

    def __init__(self, nc: int, strides: Sequence[int], reg_max: int = 16):
        super().__init__()
        self.nc, self.no, self.nl = nc, nc + 5, len(strides)
        self.stride = torch.tensor(list(strides))
        self.reg_max, self.use_dfl = reg_max, True
        self.grid = [torch.zeros(1)] * self.nl  # touched by the reference Model._apply (yolo.py:211-215)
        self.eval()  # inference-only drop-in

    @torch.no_grad()
    def forward(self, x, val_loss: bool = False):
        if val_loss or self.training:
            raise NotImplementedError("training branch (yolo.py:333-354) is out of scope")
        def to_nhwc(t):
            n, c, h, w = t.shape
            dst = NHWC.empty(n, h, w, c, t.device)
            ops.nchw_to_nhwc(t.contiguous() if t.dtype in (torch.float32, torch.float16) else t.float().contiguous(), dst)
            return dst

        with torch.cuda.device(x[0][0].device):
            cls = [to_nhwc(lv[1]) for lv in x]
            reg = [to_nhwc(lv[2]) for lv in x]
        b = x[0][0].shape[0]
        anchors = sum(t.h * t.w for t in cls)
        pred = torch.empty((b, anchors, self.no), dtype=torch.float32, device=x[0][0].device)
        with torch.cuda.device(pred.device):
            ops.head_decode(cls, reg, [float(s) for s in self.stride.tolist()], self.reg_max, pred, cls_is_prob=True)
        return pred


def convert_blocks(model: torch.nn.Module, bn_eps: Optional[float] = None) -> torch.nn.Module:
    """In-place surgery on a reference `Model`: supported blocks -> B200Block, detect -> B200Detect."""
    rows = getattr(model, "yaml", None)
    if rows is None:
        raise TypeError("convert_blocks() expects a yaml-built reference Model")
    nc = model.detect.nc
    if bn_eps is None:
        bn_eps = next((m.eps for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)), 1e-3)
    graph = build_graph(rows, nc)
    folded = fold_state_dict(graph, model.state_dict(), bn_eps)
    for lay in graph.layers:
        if lay.kind in SUPPORTED:
            old = model.backbone[lay.i]
            blk = B200Block(graph, folded, lay.i)
            blk.i, blk.f, blk.np = old.i, old.f, getattr(old, "np", 0)
            model.backbone[lay.i] = blk
    reg_max = graph.layers[graph.head_layers[0]].reg_max
    model.detect = B200Detect(nc, graph.strides, reg_max)
    return model.eval()
