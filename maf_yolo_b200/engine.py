"""Static execution plan for the MAF-YOLO forward -> decode path on one B200.

`plan_graph()` turns the resolved topology into a flat list of kernel launches over NHWC fp16
buffers (no per-forward Python graph walking, no torch ops):

  * Concat / split never move data: producers write channel slices of a shared buffer
    (RepHDW, MPRep, SPPF — yolov6/layers/common.py:938-946,787-792,123-129) and the MAFPN fusion
    Concats (configs/yaml/MAF-YOLO-n.yaml:18-42) become multi-source K loops of the consuming 1x1 GEMM.
  * nn.Upsample is fused into the epilogue of the conv that produces its input (dual store).
  * buffers get arena offsets from a liveness scan, so the working set is small and L2-friendly.

`Engine` binds a plan to device memory + packed weights and replays it — eagerly or as one CUDA
graph.  The input tensor is read directly in the reference's format (NCHW fp32/fp16/uint8) by the
stem kernel; the result is the reference's `[B, A, 5+nc]` fp32 prediction tensor.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .fold import Folded
from .ops import NHWC
from .topology import Graph, Layer


def _ld(c: int) -> int:
    """Channel stride of an arena buffer: a multiple of 16 fp16 = 32 B, so every pixel row starts on a
    32-byte sector boundary and the GEMM epilogue can use 256-bit stores (measured -10 % GEMM time)."""
    return (c + 15) // 16 * 16


class _Buf:
    __slots__ = ("h", "w", "c", "ld", "first", "last", "offset", "name")

    def __init__(self, h, w, c, name):
        self.h, self.w, self.c, self.ld, self.name = h, w, c, _ld(c), name
        self.first, self.last, self.offset = None, None, None

    def touch(self, idx):
        self.first = idx if self.first is None else min(self.first, idx)
        self.last = idx if self.last is None else max(self.last, idx)

    def nbytes(self, batch):
        return batch * self.h * self.w * self.ld * 2


@dataclass(frozen=True)
class _View:
    buf: _Buf
    c_off: int
    c: int

    def slice(self, off, c):
        assert off + c <= self.c
        return _View(self.buf, self.c_off + off, c)


@dataclass
class _Op:
    kind: str
    name: str
    reads: List[_View]
    writes: List[_View]
    weight: Optional[str] = None
    act: str = "none"
    k: int = 0
    weight2: Optional[str] = None             # fused depth-wise + 1x1 ("dwpw"): the 1x1's folded weight
    weight3: Optional[str] = None             # fused bottleneck ("bneck"): weight = expand 1x1, weight3 = depth-wise, weight2 = project 1x1
    act2: str = "none"
    wslice: Optional[Tuple[int, int]] = None  # channel range of the folded weight this op uses (split depth-wise convs)
    level: int = -1                           # head_pred: pyramid level; anchor_off = first anchor of the level
    anchor_off: int = 0
    # dependencies through memory outside the arena (pred / boxes / NMS workspace), by name
    tag_reads: Tuple[str, ...] = ()
    tag_writes: Tuple[str, ...] = ()
    flops_per_image: int = 0
    bytes_per_image: int = 0   # algorithmic: activations read once + written once (fp16), weights excluded


def pad_fill_channels(op: "_Op", act: str) -> int:
    """How many padding channels behind the op's output it may fill with zeros (Engine._pad_fill): the op must be the
    only writer of the LAST channels of its buffer, those must end inside a 32-byte sector, the channel stride must be
    wider than the buffer's channels, and act(0) must be 0 (not sigmoid).  MAFB200_PAD_FILL=0 turns it off."""
    v = op.writes[0]
    b = v.buf
    if len(op.writes) != 1 or v.c_off + v.c != b.c or b.ld <= b.c or act == "sigmoid":
        return 0
    if ((v.c_off + v.c) * 2) % 32 == 0 or os.environ.get("MAFB200_PAD_FILL", "1") == "0":
        return 0
    return b.ld - b.c


class Plan:
    """Kernel-launch schedule + buffer liveness for one (graph, H, W); batch-size independent."""

    def __init__(self, graph: Graph, height: int, width: int, only_layer: Optional[int] = None,
                 head_sigmoid: bool = False, k7: Optional[bool] = None):
        """only_layer: plan a single yaml layer (block-level drop-in); its sources become input buffers
        (`self.inputs`), (height, width) is then the spatial size of those sources."""
        if only_layer is None:
            assert height % 32 == 0 and width % 32 == 0, "input size must be a multiple of the largest stride (32)"
        self.graph, self.height, self.width = graph, height, width
        self.only_layer, self.head_sigmoid = only_layer, head_sigmoid
        # K7: cls_pred / reg_pred finish the detect path in their GEMM epilogues (mafb200_head_pred); no cls / reg maps,
        # no decode kernel.  Needs the whole graph (anchor offsets), reg_max 16 and nc <= 128.  MAFB200_K7=0: the
        # round-1 form (fp16 cls / reg maps + head_decode kernel).
        if k7 is None:
            k7 = os.environ.get("MAFB200_K7", "1") != "0"
        heads = [graph.layers[i] for i in graph.head_layers]
        self.k7 = bool(k7 and only_layer is None and not head_sigmoid and heads and
                       all(h.reg_max == 16 for h in heads) and graph.nc <= 128)
        self.bufs: List[_Buf] = []
        self.ops: List[_Op] = []
        self.level_views: List[Tuple[_View, _View, _View]] = []  # (stem, cls_logits, reg) per head
        self.layer_out: Dict[int, object] = {}  # yaml layer index -> output view(s)
        self.inputs: List[_View] = []
        self.anchors = 0
        self._plan()

    # ---- helpers --------------------------------------------------------------------------------
    def _buf(self, h, w, c, name) -> _View:
        b = _Buf(h, w, c, name)
        self.bufs.append(b)
        return _View(b, 0, c)

    def _emit(self, kind, name, reads, writes, **kw) -> _Op:
        idx = len(self.ops)
        mid = kw.pop("mid", None)  # accounting only (fused bottleneck)
        for v in list(reads) + list(writes):
            v.buf.touch(idx)
        op = _Op(kind, name, list(reads), list(writes), **kw)
        px_out = writes[0].buf.h * writes[0].buf.w if writes else 0
        op.bytes_per_image = sum(2 * v.c * v.buf.h * v.buf.w for v in reads) + sum(
            2 * v.c * v.buf.h * v.buf.w for v in writes)
        if kind == "conv1x1":
            op.flops_per_image = 2 * sum(v.c for v in reads) * writes[0].c * px_out
        elif kind == "conv3x3s2":
            op.flops_per_image = 2 * 9 * reads[0].c * writes[0].c * px_out
        elif kind == "dwconv":
            op.flops_per_image = 2 * kw["k"] * kw["k"] * writes[0].c * px_out
        elif kind == "poolpw":
            op.flops_per_image = 2 * reads[0].c * writes[0].c * px_out
        elif kind == "dwpw":
            op.flops_per_image = 2 * kw["k"] * kw["k"] * reads[0].c * px_out + 2 * reads[0].c * writes[0].c * px_out
        elif kind == "bneck":  # expand 1x1 + depth-wise + project 1x1 (no halo recompute counted: algorithmic)
            op.flops_per_image = 2 * px_out * (reads[0].c * mid + kw["k"] * kw["k"] * mid + mid * writes[0].c)
        self.ops.append(op)
        return op

    def _emit_dw(self, name: str, src: _View, dst: _View, weight: str, act: str, k: int):
        self._emit("dwconv", name, [src], [dst], weight=weight, act=act, k=k)

    @staticmethod
    def _dwpw_ok(c: int, cout: int, k: int) -> bool:
        """Can `depth-wise k x k -> 1x1` run as ONE kernel (mafb200_dwconv_conv1x1)?  k <= 5, one column tile, and the
        A tile + W2 panel + halo tile must fit 113 KB so that two CTAs stay resident per SM."""
        if os.environ.get("MAFB200_DWPW", "1") == "0" or k > 5 or c % 8 or cout % 8 or cout > 128:
            return False
        tile_n = (cout + 15) // 16 * 16
        halo = (10 + k - 1) * (20 + k - 1) * 64 * 2
        return 1024 + 32768 + 2 * tile_n * 128 + (halo + 127) // 128 * 128 + 64 + tile_n * 4 <= 113 * 1024

    @staticmethod
    def _bneck_ok(c_: int, mid: int, k: int) -> bool:
        """Should the whole DepthBottleneckUni run as ONE kernel (mafb200_bottleneck, K4)?  Opt-in (MAFB200_BNECK=1):
        the kernel is parity-green and removes the 3c_-wide buffer from the plan (160 -> 131.5 MB/img for N), but on B200
        it measured 0.89x the speed of the two-kernel form it replaces (expand GEMM, then depth-wise + project:
        profiles/r02_d_bneck_*.txt — its tap warps are issue/latency-bound at 2 warps per scheduler), so the default
        plan keeps the two-kernel form."""
        if os.environ.get("MAFB200_BNECK", "0") != "1":
            return False
        return ops.bottleneck_supported(c_, mid, (c_ + 15) // 16 * 16, k)

    # ---- the schedule ----------------------------------------------------------------------------
    def _plan(self):
        g = self.graph
        H, W = self.height, self.width
        up_of: Dict[int, int] = {l.frm[0]: l.i for l in g.layers if l.kind == "upsample"}
        out: Dict[int, object] = {}  # layer index -> _View | list[_View] | tuple (head)
        layers = g.layers
        if self.only_layer is not None:
            lay = g.layers[self.only_layer]
            layers = [lay]
            up_of = {}
            for s_idx, c in zip(lay.frm, lay.c_in):
                if s_idx >= 0:
                    v = self._buf(H, W, c, f"in{s_idx}")
                    out[s_idx] = v
                    self.inputs.append(v)

        def size(l: Layer):
            if self.only_layer is not None:  # (H, W) is the block's input size
                down = 2 if l.kind in ("repvgg", "convw", "mprep") else 1
                return H // down, W // down
            return H // l.stride_total, W // l.stride_total

        def srcs_of(l: Layer) -> List[_View]:
            res: List[_View] = []
            for s in l.frm:
                v = out[s]
                res.extend(v if isinstance(v, list) else [v])
            return res

        for l in layers:
            i = str(l.i)
            if l.kind == "repvgg":
                h, w = size(l)
                dst = self._buf(h, w, l.c_out, f"L{i}")
                if l.frm[0] < 0:
                    op = self._emit("stem", f"L{i}.stem3x3s2", [], [dst], weight=i, act="relu")
                    op.flops_per_image = 2 * 27 * l.c_out * h * w
                    op.bytes_per_image += 3 * H * W * 4  # the fp32 NCHW image
                else:
                    self._emit("conv3x3s2", f"L{i}.repvgg3x3s2", [out[l.frm[0]]], [dst], weight=i, act="relu")
                out[l.i] = dst
            elif l.kind == "convw":
                h, w = size(l)
                dst = self._buf(h, w, l.c_out, f"L{i}")
                self._emit("conv3x3s2", f"L{i}.conv3x3s2", [out[l.frm[0]]], [dst], weight=i + ".block", act="silu")
                out[l.i] = dst
            elif l.kind == "rephdw":
                h, w = size(l)
                c_, mid = l.c_hidden, l.expand * l.c_hidden
                cat = self._buf(h, w, (2 + l.depth) * c_, f"L{i}.cat")
                self._emit("conv1x1", f"L{i}.conv1", srcs_of(l), [cat.slice(0, 2 * c_)], weight=i + ".conv1", act="silu")
                for j in range(l.depth):
                    if self._bneck_ok(c_, mid, l.k):  # K4: expand + depth-wise + project in one kernel, no 3c_ buffer
                        self._emit("bneck", f"L{i}.m{j}.bottleneck(k{l.k})", [cat.slice((1 + j) * c_, c_)],
                                   [cat.slice((2 + j) * c_, c_)], weight=f"{i}.m.{j}.conv1", weight3=f"{i}.m.{j}.dw",
                                   weight2=f"{i}.m.{j}.one_conv", act="silu", act2="silu", k=l.k, mid=mid)
                        continue
                    t1 = self._buf(h, w, mid, f"L{i}.m{j}.expand")
                    self._emit("conv1x1", f"L{i}.m{j}.conv1", [cat.slice((1 + j) * c_, c_)], [t1],
                               weight=f"{i}.m.{j}.conv1", act="silu")
                    if self._dwpw_ok(mid, c_, l.k):  # depth-wise + one_conv in one kernel: no 3c_-wide round trip
                        self._emit("dwpw", f"L{i}.m{j}.dw{l.k}+one_conv", [t1], [cat.slice((2 + j) * c_, c_)],
                                   weight=f"{i}.m.{j}.dw", act="silu", k=l.k, weight2=f"{i}.m.{j}.one_conv", act2="silu")
                        continue
                    t2 = self._buf(h, w, mid, f"L{i}.m{j}.dw")
                    self._emit_dw(f"L{i}.m{j}.dw{l.k}", t1, t2, f"{i}.m.{j}.dw", "silu", l.k)
                    self._emit("conv1x1", f"L{i}.m{j}.one_conv", [t2], [cat.slice((2 + j) * c_, c_)],
                               weight=f"{i}.m.{j}.one_conv", act="silu")
                dst = self._buf(h, w, l.c_out, f"L{i}")
                writes = [dst]
                if l.i in up_of:  # fuse the following nn.Upsample into this conv's epilogue
                    up = self._buf(2 * h, 2 * w, l.c_out, f"L{up_of[l.i]}.up")
                    writes.append(up)
                    out[up_of[l.i]] = up
                self._emit("conv1x1", f"L{i}.conv2", [cat], writes, weight=i + ".conv2", act="silu")
                out[l.i] = dst
            elif l.kind == "mprep":
                h, w = size(l)
                src = out[l.frm[0]]
                dst = self._buf(h, w, l.c_out, f"L{i}")
                half = l.c_out // 2
                if (os.environ.get("MAFB200_POOLPW", "1") != "0" and l.c_in[0] <= 256 and l.c_in[0] % 8 == 0 and
                        half <= 128 and half % 8 == 0):  # max pool + conv1 in one kernel: no pooled map in HBM
                    self._emit("poolpw", f"L{i}.maxpool+conv1", [src], [dst.slice(0, half)], weight=i + ".conv1", act="silu")
                else:
                    pooled = self._buf(h, w, l.c_in[0], f"L{i}.pool")
                    self._emit("maxpool2x2", f"L{i}.maxpool", [src], [pooled])
                    self._emit("conv1x1", f"L{i}.conv1", [pooled], [dst.slice(0, half)], weight=i + ".conv1", act="silu")
                self._emit("conv3x3s2", f"L{i}.repvgg3x3s2", [src], [dst.slice(half, half)], weight=i + ".conv2",
                           act="relu")
                out[l.i] = dst
            elif l.kind == "sppf":
                h, w = size(l)
                c_ = l.c_hidden
                cat = self._buf(h, w, 4 * c_, f"L{i}.cat")
                self._emit("conv1x1", f"L{i}.cv1", srcs_of(l), [cat.slice(0, c_)], weight=i + ".cv1", act="silu")
                self._emit("sppf_pool", f"L{i}.pool5x3", [cat.slice(0, c_)],
                           [cat.slice(c_, c_), cat.slice(2 * c_, c_), cat.slice(3 * c_, c_)])
                dst = self._buf(h, w, l.c_out, f"L{i}")
                self._emit("conv1x1", f"L{i}.cv2", [cat], [dst], weight=i + ".cv2", act="silu")
                out[l.i] = dst
            elif l.kind == "concat":
                views = srcs_of(l)
                if len(views) > 4:
                    raise NotImplementedError("a fusion stage with more than 4 inputs is outside MAF-YOLO's topology")
                out[l.i] = views
            elif l.kind == "upsample":
                if l.i not in out:  # producer could not dual-store: standalone kernel
                    src = out[l.frm[0]]
                    h, w = size(l)
                    dst = self._buf(h, w, l.c_out, f"L{i}.up")
                    self._emit("upsample2x", f"L{i}.upsample", [src], [dst])
                    out[l.i] = dst
            elif l.kind == "head":
                h, w = size(l)
                c = l.c_out
                stem = self._buf(h, w, c, f"L{i}.stem")
                self._emit("conv1x1", f"L{i}.stem", srcs_of(l), [stem], weight=i + ".stem", act="silu")
                res = {}
                lvl = len(self.level_views)
                if self.k7 and lvl == 0:  # zero the NMS candidate counters once per forward, ahead of every cls_pred
                    self._emit("detect_reset", "detect.reset", [], [], tag_writes=("ncand",))
                for br, cout in (("cls", g.nc), ("reg", 4 * (l.reg_max + 1))):
                    f2 = self._buf(h, w, c, f"L{i}.{br}_s")
                    if self._dwpw_ok(c, c, l.k):
                        self._emit("dwpw", f"L{i}.{br}_dw{l.k}+{br}_s", [stem], [f2], weight=f"{i}.{br}_dw", act="none",
                                   k=l.k, weight2=f"{i}.{br}_s", act2="silu")
                    else:
                        t = self._buf(h, w, c, f"L{i}.{br}_dw")
                        self._emit_dw(f"L{i}.{br}_dw{l.k}", stem, t, f"{i}.{br}_dw", "none", l.k)
                        self._emit("conv1x1", f"L{i}.{br}_s", [t], [f2], weight=f"{i}.{br}_s", act="silu")
                    if self.k7:
                        op = self._emit("head_pred", f"L{i}.{br}_pred+{'sigmoid' if br == 'cls' else 'dfl_decode'}", [f2], [],
                                        weight=f"{i}.{br}_pred", act=br, level=lvl, anchor_off=self.anchors,
                                        tag_reads=("ncand",) if br == "cls" else (), tag_writes=(f"out.{br}{lvl}",))
                        op.flops_per_image = 2 * c * cout * h * w
                        op.bytes_per_image += 16 * h * w if br == "reg" else 0  # serving path: boxes only (fp32 x 4)
                        res[br] = None
                        continue
                    o = self._buf(h, w, cout, f"L{i}.{br}_pred")
                    self._emit("conv1x1", f"L{i}.{br}_pred", [f2], [o], weight=f"{i}.{br}_pred",
                               act="sigmoid" if (br == "cls" and self.head_sigmoid) else "none")
                    res[br] = o
                out[l.i] = (stem, res["cls"], res["reg"])
                self.level_views.append(out[l.i])
                self.anchors += h * w
            elif l.kind == "out":
                if self.k7:
                    continue
                cls = [out[s][1] for s in l.frm]
                reg = [out[s][2] for s in l.frm]
                op = self._emit("decode", "detect.decode", cls + reg, [], tag_writes=("out",))
                op.bytes_per_image += self.anchors * (5 + g.nc) * 4
            else:
                raise NotImplementedError(l.kind)
        self.layer_out = out

    # ---- arena ------------------------------------------------------------------------------------
    def assign_offsets(self, batch: int, align: int = 1024, reuse: bool = True) -> int:
        """Greedy interval allocation: buffers whose live ranges do not overlap share memory.
        reuse=False gives every buffer its own range (debugging: intermediates stay readable)."""
        live: List[_Buf] = []
        total = 0
        for b in sorted(self.bufs, key=lambda b: (b.first, -b.nbytes(batch))):
            if reuse:
                live = [x for x in live if x.last >= b.first]
            size = (b.nbytes(batch) + align - 1) // align * align
            off = 0
            for x in sorted(live, key=lambda x: x.offset):
                xs = (x.nbytes(batch) + align - 1) // align * align
                if off + size <= x.offset:
                    break
                off = max(off, x.offset + xs)
            b.offset = off
            live.append(b)
            total = max(total, off + size)
        return total

    # ---- accounting (DESIGN.md / bench.py roofline) ---------------------------------------------------
    def flops_per_image(self) -> int:
        return sum(o.flops_per_image for o in self.ops)

    def bytes_per_image(self) -> int:
        return sum(o.bytes_per_image for o in self.ops)

    def summary(self) -> List[dict]:
        return [dict(name=o.name, kind=o.kind, flops=o.flops_per_image, bytes=o.bytes_per_image) for o in self.ops]


class Engine:
    """A plan bound to one device, one batch size and one set of folded weights."""

    def __init__(self, graph: Graph, folded: Folded, batch: int, height: int = 640, width: int = 640,
                 device: Optional[torch.device] = None, use_cuda_graph: bool = True, reuse_buffers: Optional[bool] = None,
                 n_streams: int = 4, k7: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("maf_yolo_b200.Engine needs a B200 GPU: the hot path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self.graph, self.batch, self.height, self.width = graph, batch, height, width
        self.plan = Plan(graph, height, width, k7=k7)
        self.use_cuda_graph = use_cuda_graph
        # Branch-level concurrency (captured into the CUDA graph): independent chains of the yaml graph —
        # the three heads and their cls / reg towers, the MAFPN down-sampling convs — run on side streams.
        # Aliased arena buffers would serialise unrelated branches, so multi-stream plans do not reuse memory.
        self.n_streams = max(1, n_streams) if use_cuda_graph else 1
        if reuse_buffers is None:
            reuse_buffers = self.n_streams == 1
        self.reuse_buffers = reuse_buffers
        with torch.cuda.device(self.device):
            nbytes = self.plan.assign_offsets(batch, reuse=reuse_buffers)
            self.arena = torch.zeros(nbytes // 2, dtype=torch.float16, device=self.device)  # padding channels finite
            # two prediction buffers, used alternately: the NMS of call i may still read preds[i % 2] on another
            # stream while call i+1 runs (B200DetectModel.detect_async); sequential callers never notice
            self.preds = [torch.empty((batch, self.plan.anchors, 5 + graph.nc), dtype=torch.float32, device=self.device)
                          for _ in range(2)]
            self.pred = self.preds[0]
            self._flip = 0
            self.last_index = 0
            # event of the last side-stream reader (detect_async's NMS) of each prediction buffer: the next
            # forward that overwrites the buffer waits for it, whichever API the caller mixes
            self.reader_done: List[Optional[torch.cuda.Event]] = [None, None]
            # completion of the last forward that detect_async issued on ITS stream: a later forward of this
            # engine from another stream must not touch the arena before it
            self.last_async_forward: Optional[torch.cuda.Event] = None
            self._views: Dict[Tuple[int, int, int], NHWC] = {}
            self._weights: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
            self._x: Optional[torch.Tensor] = None
            self._calls: List[Callable[[], None]] = [self._bind(op, folded) for op in self.plan.ops]
        # captured graphs, keyed by (prediction-buffer index, detect configuration or None)
        self._graphs: Dict[Tuple[int, object], torch.cuda.CUDAGraph] = {}
        # whole-forward graphs (stem included) per input address, and how often an address has been seen
        self._full_graphs: Dict[object, torch.cuda.CUDAGraph] = {}
        self._ptr_seen: Dict[object, int] = {}
        # detect mode (B200DetectModel.detect_async): the decode kernel also does the NMS threshold / compaction pass
        # and the prediction tensor is not materialised; buffers are allocated on first use
        self._detect: Optional[Tuple[float, bool, Optional[Tuple[int, ...]]]] = None
        self._detect_filters: Dict[Tuple[int, ...], torch.Tensor] = {}
        self.boxes: Optional[List[torch.Tensor]] = None
        self.nms_ws: Optional[List[torch.Tensor]] = None
        self.train_out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None  # (pred_scores, pred_distri), mode "train"
        # K7 detect mode: thresholds / class filter live in a small DEVICE struct the cls_pred epilogues read when they
        # run, so one captured graph serves every (conf_thres, multi_label, classes); host copies are immutable, cached
        self.detect_cfg_dev: Optional[torch.Tensor] = None
        self._detect_cfg_host: Dict[object, torch.Tensor] = {}
        self._detect_cfg_current = None
        self.launches_per_forward = sum(1 for o in self.plan.ops if o.kind != "detect_reset")  # the reset is a memset
        self._schedule = self._make_schedule() if self.n_streams > 1 else None
        self._side_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams - 1)]

    # ---- binding ----------------------------------------------------------------------------------
    def view(self, v: _View) -> NHWC:
        key = (id(v.buf), v.c_off, v.c)
        hit = self._views.get(key)
        if hit is None:
            b = v.buf
            numel = self.batch * b.h * b.w * b.ld
            t = self.arena[b.offset // 2: b.offset // 2 + numel].view(self.batch, b.h, b.w, b.ld)
            hit = NHWC(t, v.c_off, v.c)
            self._views[key] = hit
        return hit

    def _pad_fill(self, op: _Op, act: str, wt: torch.Tensor, bs: torch.Tensor):
        """Whole-sector output rows.  When the op writes the LAST channels of a buffer whose channel stride is wider
        (24 of 32, 72 of 80) and they end in the middle of a 32-byte sector, every pixel's last store is half a sector:
        a slow 16-byte store and a DRAM read-modify-write (ncu: the stem read 96 MB more than its input).  The padding
        channels belong to nobody, so give the op zero filters for them: act(0) = 0 lands there at no cost (the column
        tile is already that wide).  Returns (output view, weight, bias), extended or as they were."""
        v = op.writes[0]
        b = v.buf
        extra = pad_fill_channels(op, act)
        if extra == 0:
            return self.view(v), wt, bs
        wt = torch.cat([wt, wt.new_zeros((extra,) + tuple(wt.shape[1:]))])
        bs = torch.cat([bs, bs.new_zeros(extra)])
        return self.view(_View(b, v.c_off, v.c + extra)), wt, bs

    def _bind(self, op: _Op, folded: Folded) -> Callable[[], None]:
        dev = self.device
        reads = [self.view(v) for v in op.reads]
        writes = [self.view(v) for v in op.writes]
        if op.kind == "stem":
            dst, wt, bs = self._pad_fill(op, op.act, *folded[op.weight])
            w, b = ops.pack_stem(wt, bs, device=dev)
            self._weights[op.name] = (w, b)
            return lambda: ops.stem_conv3x3s2(self._x, w, b, op.act, dst)
        if op.kind == "conv1x1":
            wt, bs = folded[op.weight]
            if len(writes) == 1:
                writes[0], wt, bs = self._pad_fill(op, op.act, wt, bs)
            w, b = ops.pack_conv1x1(wt.reshape(wt.shape[0], -1), bs, [r.c for r in reads], device=dev)
            self._weights[op.name] = (w, b)
            up = writes[1] if len(writes) > 1 else None
            return lambda: ops.conv1x1(reads, w, b, op.act, writes[0], up)
        if op.kind == "conv3x3s2":
            src = reads[0]
            if (2 * src.ld == 64 and src.c_off == 0 and src.w % 2 == 0 and
                    os.environ.get("MAFB200_CONV3_PAIR", "1") != "0"):
                # narrow map (N / S layer 1): pixel-pair im2col, 6 boxes per tile instead of 9.  Its zero weights meet
                # the padding channels of the source buffer, which therefore must be finite: the arena is zero-filled.
                w, b = ops.pack_conv3x3_pair(*folded[op.weight], ld=src.ld, device=dev)
                self._weights[op.name] = (w, b)
                return lambda: ops.conv3x3s2_pair(src, w, b, op.act, writes[0])
            w, b = ops.pack_conv3x3(*folded[op.weight], device=dev)
            self._weights[op.name] = (w, b)
            return lambda: ops.conv3x3s2(reads[0], w, b, op.act, writes[0])
        if op.kind == "bneck":
            dst, wt2, bs2 = self._pad_fill(op, op.act2, *folded[op.weight2])
            packed = ops.pack_bottleneck(*folded[op.weight], *folded[op.weight3], wt2, bs2, device=dev)
            self._weights[op.name] = packed
            return lambda: ops.bottleneck(reads[0], packed, dst)
        if op.kind == "dwpw":
            dw_w, dw_b = ops.pack_dw(*folded[op.weight], device=dev)
            dst, wt2, bs2 = self._pad_fill(op, op.act2, *folded[op.weight2])
            pw_w, pw_b = ops.pack_conv1x1(wt2.reshape(wt2.shape[0], -1), bs2, [reads[0].c], device=dev)
            self._weights[op.name] = (dw_w, dw_b, pw_w, pw_b)
            return lambda: ops.dwconv_conv1x1(reads[0], dw_w, dw_b, op.k, op.act, pw_w, pw_b, op.act2, dst)
        if op.kind == "dwconv":
            wt, bs = folded[op.weight]
            if op.wslice is not None:
                wt, bs = wt[op.wslice[0]:op.wslice[1]], bs[op.wslice[0]:op.wslice[1]]
            w, b = ops.pack_dw(wt, bs, device=dev)
            self._weights[op.name] = (w, b)
            return lambda: ops.dwconv(reads[0], w, b, op.k, op.act, writes[0])
        if op.kind == "poolpw":
            dst, wt, bs = self._pad_fill(op, op.act, *folded[op.weight])
            w, b = ops.pack_conv1x1(wt.reshape(wt.shape[0], -1), bs, [reads[0].c], device=dev)
            self._weights[op.name] = (w, b)
            return lambda: ops.maxpool2x2_conv1x1(reads[0], w, b, op.act, dst)
        if op.kind == "maxpool2x2":
            return lambda: ops.maxpool2x2(reads[0], writes[0])
        if op.kind == "sppf_pool":
            return lambda: ops.sppf_pool(reads[0], writes[0], writes[1], writes[2])
        if op.kind == "upsample2x":
            return lambda: ops.upsample2x(reads[0], writes[0])
        if op.kind == "detect_reset":
            def reset():
                if self._detect is not None and self._detect != "train":
                    ops.detect_reset(self.nms_ws[self.last_index], self.batch)

            return reset
        if op.kind == "head_pred":
            wt, bs = folded[op.weight]
            nc, total = self.graph.nc, self.plan.anchors
            stride = float(self.graph.strides[op.level])
            if op.act == "cls":
                w, b = ops.pack_conv1x1(wt.reshape(wt.shape[0], -1), bs, [reads[0].c], device=dev)
            else:
                w, b = ops.pack_head_reg(wt, bs, device=dev)
            self._weights[op.name] = (w, b)

            def head_pred():
                if self._detect == "train":  # train-form outputs (yolo.py:333-354) for the validation loss
                    ops.head_pred(reads[0], w, b, op.act + "_train", op.anchor_off, total, stride, nc,
                                  pred=self.train_out[0 if op.act == "cls" else 1])
                elif self._detect is None:
                    ops.head_pred(reads[0], w, b, op.act, op.anchor_off, total, stride, nc, pred=self.pred)
                elif op.act == "cls":
                    ops.head_pred(reads[0], w, b, "cls", op.anchor_off, total, stride, nc, detect_cfg=self.detect_cfg_dev,
                                  workspace=self.nms_ws[self.last_index])
                else:
                    ops.head_pred(reads[0], w, b, "reg", op.anchor_off, total, stride, nc, boxes=self.boxes[self.last_index])

            return head_pred
        if op.kind == "decode":
            nl = len(reads) // 2
            strides = [float(s) for s in self.graph.strides]
            reg_max = self.graph.layers[self.graph.head_layers[0]].reg_max

            def decode():
                if self._detect is None:
                    ops.head_decode(reads[:nl], reads[nl:], strides, reg_max, self.pred)
                else:
                    conf, multi_label, classes = self._detect
                    k = self.last_index
                    ops.head_decode_detect(reads[:nl], reads[nl:], strides, reg_max, self.boxes[k], conf, multi_label,
                                           self._detect_filters[classes] if classes is not None else None, self.nms_ws[k])

            return decode
        raise NotImplementedError(op.kind)

    # ---- multi-stream schedule ------------------------------------------------------------------------
    def _make_schedule(self):
        """Dependencies from buffer overlap (RAW / WAR / WAW on arena byte ranges), then greedy list
        scheduling onto `n_streams` streams: an op continues the stream of one of its producers when that
        producer is still the stream's tail, otherwise takes the least-recently-used stream."""
        ops_ = self.plan.ops

        def ranges(views):  # (byte range of the buffer, buffer identity, channel range of the view)
            return [(v.buf.offset, v.buf.offset + v.buf.nbytes(self.batch), id(v.buf), v.c_off, v.c_off + v.c) for v in views]

        tags: Dict[str, int] = {}

        def tag_ranges(names):  # memory outside the arena (pred / boxes / NMS workspace): one fake byte range per name
            return [(-2 * tags.setdefault(t, len(tags)) - 2, -2 * tags[t] - 1, -1, 0, 1) for t in names]

        rd = [ranges(o.reads) + tag_ranges(o.tag_reads) for o in ops_]
        wr = [ranges(o.writes) + tag_ranges(o.tag_writes) for o in ops_]

        def hit(a, b):
            # byte ranges overlap, and — inside ONE buffer — so do the channel slices (disjoint channel slices of a
            # buffer never conflict: the two halves of a split depth-wise conv, the producers of a concat buffer)
            return any(x0 < y1 and y0 < x1 and (xb != yb or (xc0 < yc1 and yc0 < xc1))
                       for x0, x1, xb, xc0, xc1 in a for y0, y1, yb, yc0, yc1 in b)

        deps = []
        for j in range(len(ops_)):
            d = {i for i in range(j) if hit(wr[i], rd[j]) or hit(rd[i], wr[j]) or hit(wr[i], wr[j])}
            deps.append(d)
        n = self.n_streams
        tail = [-1] * n
        stream_of, waits = [], []
        for j in range(len(ops_)):
            if j == 0:
                s = 0
            else:
                cands = [t for t in range(n) if tail[t] in deps[j]]
                s = max(cands, key=lambda t: tail[t]) if cands else min(range(n), key=lambda t: tail[t])
            w = []
            for t in range(n):
                if t == s:
                    continue
                on_t = [i for i in deps[j] if stream_of[i] == t]
                if on_t:
                    w.append(max(on_t))
            stream_of.append(s)
            waits.append(w)
            tail[s] = j
        need_event = {i for w in waits for i in w}
        return stream_of, waits, need_event

    def _launch_scheduled(self, first: int):
        """Launches ops[first:] over the main stream + side streams with event dependencies."""
        stream_of, waits, need_event = self._schedule
        main = torch.cuda.current_stream(self.device)
        streams = [main] + self._side_streams
        fork = torch.cuda.Event()
        fork.record(main)
        for st in self._side_streams:
            st.wait_event(fork)
        events = {}
        for j in range(first, len(self._calls)):
            st = streams[stream_of[j]]
            for i in waits[j]:
                if i >= first:
                    st.wait_event(events[i])
            with torch.cuda.stream(st):
                self._calls[j]()
            if j in need_event:
                ev = torch.cuda.Event()
                ev.record(st)
                events[j] = ev
        for st in self._side_streams:
            main.wait_stream(st)

    # ---- execution ----------------------------------------------------------------------------------
    def _check_input(self, x: torch.Tensor):
        if tuple(x.shape) != (self.batch, 3, self.height, self.width):
            raise ValueError(f"engine was planned for input {(self.batch, 3, self.height, self.width)}, got {tuple(x.shape)}")
        if x.device != self.pred.device:
            raise ValueError(f"input is on {x.device}, engine on {self.pred.device}")
        if x.dtype not in (torch.float32, torch.float16, torch.uint8):
            raise TypeError(f"unsupported input dtype {x.dtype}")
        return x if x.is_contiguous() else x.contiguous()

    def run_eager(self, x: torch.Tensor) -> torch.Tensor:
        self._x = self._check_input(x)
        k = self._flip
        self._flip ^= 1
        self.last_index = k
        self.pred = self.preds[k]
        if self.reader_done[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.reader_done[k])
            self.reader_done[k] = None
        if self.last_async_forward is not None:
            torch.cuda.current_stream(self.device).wait_event(self.last_async_forward)
            self.last_async_forward = None
        self._upload_detect_cfg()
        for call in self._calls:
            call()
        return self.pred

    def _set_detect(self, detect) -> None:
        """detect = None (write the prediction tensor) or (conf_thres, multi_label, classes-tuple-or-None)."""
        self._detect = detect
        if detect is None:
            return
        if detect == "train":
            if not self.plan.k7:
                raise RuntimeError("train-form outputs need the K7 plan (MAFB200_K7=1, return_featmaps=False)")
            if self.train_out is None:
                with torch.cuda.device(self.device):
                    a = self.plan.anchors
                    self.train_out = (torch.empty((self.batch, a, self.graph.nc), dtype=torch.float32, device=self.device),
                                      torch.empty((self.batch, a, 68), dtype=torch.float32, device=self.device))
            return
        if self.boxes is None:
            with torch.cuda.device(self.device):
                a, nc = self.plan.anchors, self.graph.nc
                self.boxes = [torch.empty((self.batch, a, 4), dtype=torch.float32, device=self.device) for _ in range(2)]
                nbytes = ops.nms_workspace_bytes(self.batch, a, nc)
                self.nms_ws = [torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=self.device) for _ in range(2)]
        classes = detect[2]
        if self.plan.k7:
            if self.detect_cfg_dev is None:
                from ._lib import DETECT_CFG_BYTES

                self.detect_cfg_dev = torch.zeros(DETECT_CFG_BYTES, dtype=torch.uint8, device=self.device)
            if detect not in self._detect_cfg_host:
                if len(self._detect_cfg_host) >= 64:  # bounded: drop the oldest configuration
                    self._detect_cfg_host.pop(next(iter(self._detect_cfg_host)))
                self._detect_cfg_host[detect] = ops.detect_cfg_host(detect[0], detect[1], self.graph.nc, classes)
            return
        if classes is not None and classes not in self._detect_filters:
            filt = torch.zeros(self.graph.nc, dtype=torch.uint8)
            for c in classes:
                if 0 <= int(c) < self.graph.nc:
                    filt[int(c)] = 1
            self._detect_filters[classes] = filt.to(self.device)

    def _upload_detect_cfg(self) -> None:
        """K7 detect mode: stream-ordered copy of the configuration struct (only when it changed).  Called after the
        waits that order this forward behind the previous one, so no launch in flight still reads the old values."""
        if self._detect is not None and self._detect != "train" and self.plan.k7 and self._detect_cfg_current != self._detect:
            self.detect_cfg_dev.copy_(self._detect_cfg_host[self._detect], non_blocking=True)
            self._detect_cfg_current = self._detect

    def _graph_key(self, k: int, detect):
        """Captured graphs: per prediction buffer and — K7 — per MODE only (thresholds are read from device memory);
        without K7 the thresholds are launch arguments, so per configuration, bounded to the 8 most recent."""
        if self.plan.k7:
            return (k, "train" if detect == "train" else detect is not None)
        key = (k, detect)
        if key not in self._graphs and len(self._graphs) >= 8:
            self._graphs.pop(next(iter(self._graphs)))
        return key

    def forward(self, x: torch.Tensor, detect=None) -> torch.Tensor:
        """x: NCHW fp32/fp16 in [0,1] or uint8 -> pred [B, A, 5+nc] fp32 (engine-owned buffer); with `detect`
        (see _set_detect) -> boxes [B, A, 4], the NMS candidates being left in self.nms_ws[self.last_index]."""
        self._set_detect(detect)
        if not self.use_cuda_graph:
            self.run_eager(x)
            return self._result(detect, self.last_index)
        x = self._check_input(x)
        self._x = x
        k = self._flip
        self._flip ^= 1
        self.last_index = k
        self.pred = self.preds[k]  # the decode launch reads self.pred when it is issued / captured
        if self.reader_done[k] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.reader_done[k])
            self.reader_done[k] = None
        if self.last_async_forward is not None:
            torch.cuda.current_stream(self.device).wait_event(self.last_async_forward)
            self.last_async_forward = None
        self._upload_detect_cfg()
        # The stem kernel reads the caller's tensor.  Serving loops feed a few staging buffers over and over, so once an
        # input ADDRESS has been seen twice the whole forward — stem included — is captured for that address and a step
        # is ONE graph launch (bounded: the 8 most recent addresses).  For a new address the stem is launched eagerly
        # and everything behind it, which only touches engine-owned memory, is replayed as one graph.
        gkey = self._graph_key(k, detect)
        ptr = x.data_ptr()
        fkey = (gkey, ptr, x.dtype)
        full = self._full_graphs.get(fkey)
        if full is not None:
            full.replay()
            return self._result(detect, k)
        if gkey not in self._graphs:
            self._calls[0]()
            for call in self._calls[1:]:  # warm-up outside capture (sets func attributes, loads modules)
                call()
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if self._schedule is not None:
                    self._launch_scheduled(1)
                else:
                    for call in self._calls[1:]:
                        call()
            self._graphs[gkey] = g
        seen = self._ptr_seen.get(fkey, 0) + 1
        self._ptr_seen[fkey] = seen
        if seen >= 2 and os.environ.get("MAFB200_FULL_GRAPH", "1") != "0":
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._calls[0]()
                if self._schedule is not None:
                    self._launch_scheduled(1)
                else:
                    for call in self._calls[1:]:
                        call()
            if len(self._full_graphs) >= 8:
                self._full_graphs.pop(next(iter(self._full_graphs)))
            if len(self._ptr_seen) > 64:
                self._ptr_seen.clear()
            self._full_graphs[fkey] = g
            g.replay()
            return self._result(detect, k)
        self._calls[0]()
        self._graphs[gkey].replay()
        return self._result(detect, k)

    def _result(self, detect, k: int):
        if detect is None:
            return self.pred
        return self.train_out if detect == "train" else self.boxes[k]

    __call__ = forward

    def head_outputs(self) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """(stem, cls_logits, reg) of each level as NCHW fp32 copies (debug / featmaps)."""
        return [tuple(self.view(v).to_nchw() for v in lv) for lv in self.plan.level_views]

    def layer_output(self, i: int) -> torch.Tensor:
        """NCHW fp32 copy of yaml layer i's output (only meaningful with reuse_buffers=False)."""
        v = self.plan.layer_out[i]
        if isinstance(v, list):
            return torch.cat([self.view(x).to_nchw() for x in v], 1)
        if isinstance(v, tuple):
            raise ValueError("head layers have three outputs: use head_outputs()")
        return self.view(v).to_nchw()

    def arena_bytes(self) -> int:
        return self.arena.numel() * 2
