"""Thin Python wrappers over the C ABI: one function per entry point of include/mafb200.h.

PyTorch is used here only for device memory and the CUDA stream; every op is a call into
libmafb200.so.  Activations are NHWC fp16 buffers; `NHWC` is a (possibly channel-sliced) view.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_CODES, MAF_F16, MAF_F32, MAF_U8, MafTensor, check, lib

_TORCH2MAF = {torch.float16: MAF_F16, torch.float32: MAF_F32, torch.uint8: MAF_U8}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _act(act) -> int:
    return ACT_CODES[act] if isinstance(act, str) else int(act)


class NHWC:
    """View of `c` channels starting at `c_off` inside an NHWC fp16 buffer of shape [N,H,W,ld]."""

    __slots__ = ("buf", "c_off", "c", "_maf")

    def __init__(self, buf: torch.Tensor, c_off: int = 0, c: Optional[int] = None):
        assert buf.dim() == 4 and buf.dtype == torch.float16 and buf.is_contiguous(), "NHWC fp16 contiguous buffer expected"
        self.buf = buf
        self.c_off = c_off
        self.c = buf.shape[3] - c_off if c is None else c
        assert 0 <= c_off and self.c > 0 and c_off + self.c <= buf.shape[3]
        self._maf = MafTensor(buf.data_ptr() + 2 * c_off, buf.shape[0], buf.shape[1], buf.shape[2], self.c,
                              buf.shape[3], MAF_F16)

    @staticmethod
    def empty(n: int, h: int, w: int, c: int, device="cuda", ld: Optional[int] = None) -> "NHWC":
        ld = (c + 7) // 8 * 8 if ld is None else ld
        return NHWC(torch.empty((n, h, w, ld), dtype=torch.float16, device=device), 0, c)

    @staticmethod
    def from_nchw(x: torch.Tensor, ld: Optional[int] = None) -> "NHWC":
        """Host-side helper for tests: NCHW float tensor -> NHWC fp16 buffer (torch ops, not a kernel)."""
        n, c, h, w = x.shape
        ld = (c + 7) // 8 * 8 if ld is None else ld
        buf = torch.zeros((n, h, w, ld), dtype=torch.float16, device=x.device)
        buf[..., :c] = x.permute(0, 2, 3, 1).to(torch.float16)
        return NHWC(buf, 0, c)

    n = property(lambda s: s.buf.shape[0])
    h = property(lambda s: s.buf.shape[1])
    w = property(lambda s: s.buf.shape[2])
    ld = property(lambda s: s.buf.shape[3])

    def slice(self, off: int, c: int) -> "NHWC":
        return NHWC(self.buf, self.c_off + off, c)

    def maf(self) -> MafTensor:
        return self._maf

    def ref(self):
        return C.byref(self._maf)

    def to_nchw(self) -> torch.Tensor:
        """fp32 NCHW copy (torch ops; test/boundary helper)."""
        return self.buf[..., self.c_off:self.c_off + self.c].permute(0, 3, 1, 2).float().contiguous()

    def __repr__(self):
        return f"NHWC(n={self.n}, h={self.h}, w={self.w}, c={self.c}@{self.c_off}/{self.ld})"


# ------------------------------------------------------------------------------------------------
# weight packing (layout documented in include/mafb200.h)
# ------------------------------------------------------------------------------------------------
def pack_conv1x1(weight: torch.Tensor, bias: torch.Tensor, src_channels: Sequence[int], device="cuda"):
    """weight [Cout, sum(src_channels)] (any float dtype) -> (fp16 [rows, Kp], fp32 bias [rows])."""
    cout, ktot = weight.shape
    assert ktot == sum(src_channels), (ktot, src_channels)
    n_tiles, tile_n = _lib.gemm_tiling(cout)
    rows = n_tiles * tile_n
    kp = _lib.packed_k_1x1(list(src_channels))
    wp = torch.zeros((rows, kp), dtype=torch.float16)
    ko, kpo = 0, 0
    for c in src_channels:
        wp[:cout, kpo:kpo + c] = weight[:, ko:ko + c].to(torch.float16)
        ko += c
        kpo += (c + 63) // 64 * 64
    bp = torch.zeros(rows, dtype=torch.float32)
    bp[:cout] = bias.to(torch.float32)
    return wp.to(device), bp.to(device)


def pack_conv3x3(weight: torch.Tensor, bias: torch.Tensor, device="cuda"):
    """weight [Cout, Cin, 3, 3] -> fp16 [rows, 9*round_up(Cin,64)] tap-major (ky,kx), channel fastest."""
    cout, cin, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    n_tiles, tile_n = _lib.gemm_tiling(cout)
    rows = n_tiles * tile_n
    cpad = (cin + 63) // 64 * 64
    wp = torch.zeros((rows, 9, cpad), dtype=torch.float16)
    wp[:cout, :, :cin] = weight.permute(0, 2, 3, 1).reshape(cout, 9, cin).to(torch.float16)
    bp = torch.zeros(rows, dtype=torch.float32)
    bp[:cout] = bias.to(torch.float32)
    return wp.reshape(rows, 9 * cpad).contiguous().to(device), bp.to(device)


def pack_conv3x3_pair(weight: torch.Tensor, bias: torch.Tensor, ld: int, device="cuda"):
    """weight [Cout, Cin, 3, 3] for mafb200_conv3x3s2_pair on a source with channel stride `ld` (2*ld <= 64):
    fp16 [rows, 6*64], block o = ky*2 + po over the pixel-pair view (see include/mafb200.h)."""
    cout, cin, kh, kw = weight.shape
    assert kh == 3 and kw == 3 and 2 * ld == 64 and cin <= ld
    n_tiles, tile_n = _lib.gemm_tiling(cout)
    rows = n_tiles * tile_n
    wp = torch.zeros((rows, 3, 2, 64), dtype=torch.float16)
    w16 = weight.to(torch.float16)
    wp[:cout, :, 0, ld:ld + cin] = w16[:, :, :, 0].permute(0, 2, 1)   # po = 0: second pixel of the left pair <- kx = 0
    wp[:cout, :, 1, 0:cin] = w16[:, :, :, 1].permute(0, 2, 1)         # po = 1: first pixel  <- kx = 1
    wp[:cout, :, 1, ld:ld + cin] = w16[:, :, :, 2].permute(0, 2, 1)   #         second pixel <- kx = 2
    bp = torch.zeros(rows, dtype=torch.float32)
    bp[:cout] = bias.to(torch.float32)
    return wp.reshape(rows, 6 * 64).contiguous().to(device), bp.to(device)


def pack_stem(weight: torch.Tensor, bias: torch.Tensor, device="cuda"):
    """weight [Cout, 3, 3, 3] (co, ci, ky, kx) -> fp32 [co][ky][kx][ci]."""
    return (weight.permute(0, 2, 3, 1).contiguous().to(torch.float32).to(device),
            bias.to(torch.float32).contiguous().to(device))


def pack_dw(weight: torch.Tensor, bias: torch.Tensor, device="cuda"):
    """weight [C, 1, k, k] -> fp32 [k][k][C]."""
    c, one, k, k2 = weight.shape
    assert one == 1 and k == k2
    return (weight.reshape(c, k, k).permute(1, 2, 0).contiguous().to(torch.float32).to(device),
            bias.to(torch.float32).contiguous().to(device))


def bottleneck_supported(c_in: int, mid: int, c_out: int, k: int) -> bool:
    return bool(lib().mafb200_bottleneck_supported(c_in, mid, c_out, k))


def pack_bottleneck(w1, b1, wd, bd, w2, b2, device="cuda"):
    """Folded weights of a DepthBottleneckUni -> the operands of mafb200_bottleneck (layouts in include/mafb200.h):
    w1 [mid, c_in(,1,1)], wd [mid, 1, k, k], w2 [c_out, mid(,1,1)] and their biases."""
    w1 = w1.reshape(w1.shape[0], -1)
    w2 = w2.reshape(w2.shape[0], -1)
    mid, c_in = w1.shape
    c_out, k = w2.shape[0], wd.shape[-1]
    assert w2.shape[1] == mid and wd.shape[0] == mid and c_in <= 64
    mid_pad, tile_n = (mid + 63) // 64 * 64, (c_out + 15) // 16 * 16
    w1p = torch.zeros((mid_pad, 64), dtype=torch.float16)
    w1p[:mid, :c_in] = w1.to(torch.float16)
    b1p = torch.zeros(mid_pad, dtype=torch.float32)
    b1p[:mid] = b1.to(torch.float32)
    dwp = torch.zeros((k * k, mid_pad), dtype=torch.float32)
    dwp[:, :mid] = wd.reshape(mid, k * k).t().to(torch.float32)
    bdp = torch.zeros(mid_pad, dtype=torch.float32)
    bdp[:mid] = bd.to(torch.float32)
    w2p = torch.zeros((tile_n, mid_pad), dtype=torch.float16)
    w2p[:c_out, :mid] = w2.to(torch.float16)
    b2p = torch.zeros(tile_n, dtype=torch.float32)
    b2p[:c_out] = b2.to(torch.float32)
    return tuple(t.contiguous().to(device) for t in (w1p, b1p, dwp, bdp, w2p, b2p)) + (mid, k)


def bottleneck(src: NHWC, packed, dst: NHWC) -> None:
    """K4: the whole DepthBottleneckUni (1x1 expand -> depth-wise k x k -> 1x1 project, SiLU after each) in one kernel."""
    w1p, b1p, dwp, bdp, w2p, b2p, mid, k = packed
    check(lib().mafb200_bottleneck(src.ref(), mid, w1p.data_ptr(), b1p.data_ptr(), dwp.data_ptr(), bdp.data_ptr(), k,
                                   w2p.data_ptr(), b2p.data_ptr(), dst.ref(), _stream()))


def pack_head_reg(weight: torch.Tensor, bias: torch.Tensor, device="cuda"):
    """reg_pred weight [68, C(,1,1)] / bias [68] -> packed for mafb200_head_pred(kind=REG): rows permuted so that packed
    row j < 64 is channel (j // 16) * 17 + j % 16 and row 64 + s is channel s * 17 + 16 (reg_max = 16)."""
    w = weight.reshape(weight.shape[0], -1)
    assert w.shape[0] == 68 and bias.numel() == 68, "K7 DFL epilogue is built for reg_max = 16 (68 channels)"
    perm = [(j // 16) * 17 + j % 16 for j in range(64)] + [s * 17 + 16 for s in range(4)]
    return pack_conv1x1(w[perm], bias[perm], [w.shape[1]], device=device)


def detect_cfg_host(conf_thres: float, multi_label: bool, nc: int, classes=None) -> torch.Tensor:
    """maf_detect_cfg (include/mafb200.h) as a pinned uint8 host tensor, filled by the library."""
    buf = torch.zeros(_lib.DETECT_CFG_BYTES, dtype=torch.uint8).pin_memory()
    filt = None
    if classes is not None:
        filt = torch.zeros(nc, dtype=torch.uint8)
        for c in classes:
            if 0 <= int(c) < nc:
                filt[int(c)] = 1
    check(lib().mafb200_detect_cfg_fill(buf.data_ptr(), float(conf_thres), int(bool(multi_label)), nc,
                                        filt.data_ptr() if filt is not None else None))
    return buf


# ------------------------------------------------------------------------------------------------
# ops
# ------------------------------------------------------------------------------------------------
def detect_reset(workspace: torch.Tensor, batch: int) -> None:
    check(lib().mafb200_detect_reset(workspace.data_ptr(), batch, _stream()))


def head_pred(src: NHWC, w_packed: torch.Tensor, bias: torch.Tensor, kind: str, anchor_off: int, total_anchors: int,
              stride: float, nc: int, pred: Optional[torch.Tensor] = None, boxes: Optional[torch.Tensor] = None,
              detect_cfg: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> None:
    """K7: cls_pred / reg_pred 1x1 conv of one level with sigmoid / DFL decode / candidate filter in the epilogue;
    kind "cls_train" / "reg_train": the train-form outputs (pred = pred_scores [B,A,nc] / pred_distri [B,A,68])."""
    kinds = {"cls": _lib.HEAD_CLS, "reg": _lib.HEAD_REG, "cls_train": _lib.HEAD_CLS_TRAIN, "reg_train": _lib.HEAD_REG_TRAIN}
    check(lib().mafb200_head_pred(src.ref(), w_packed.data_ptr(), bias.data_ptr(),
                                  kinds[kind], anchor_off, total_anchors,
                                  float(stride), nc, pred.data_ptr() if pred is not None else None,
                                  boxes.data_ptr() if boxes is not None else None,
                                  detect_cfg.data_ptr() if detect_cfg is not None else None,
                                  workspace.data_ptr() if workspace is not None else None,
                                  workspace.numel() * workspace.element_size() if workspace is not None else 0, _stream()))



def conv1x1(srcs: Sequence[NHWC], w_packed: torch.Tensor, bias: torch.Tensor, act, dst: NHWC,
            dst_up2x: Optional[NHWC] = None) -> None:
    arr = (MafTensor * len(srcs))(*[s.maf() for s in srcs])
    check(lib().mafb200_conv1x1(arr, len(srcs), w_packed.data_ptr(), bias.data_ptr(), _act(act), dst.ref(),
                                dst_up2x.ref() if dst_up2x is not None else None, _stream()))


def conv3x3s2(src: NHWC, w_packed: torch.Tensor, bias: torch.Tensor, act, dst: NHWC) -> None:
    check(lib().mafb200_conv3x3s2(src.ref(), w_packed.data_ptr(), bias.data_ptr(), _act(act), dst.ref(), _stream()))


def conv3x3s2_pair(src: NHWC, w_packed: torch.Tensor, bias: torch.Tensor, act, dst: NHWC) -> None:
    check(lib().mafb200_conv3x3s2_pair(src.ref(), w_packed.data_ptr(), bias.data_ptr(), _act(act), dst.ref(), _stream()))


def stem_conv3x3s2(x_nchw: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, act, dst: NHWC) -> None:
    assert x_nchw.dim() == 4 and x_nchw.shape[1] == 3 and x_nchw.is_contiguous() and x_nchw.is_cuda
    n, _, h, w = x_nchw.shape
    check(lib().mafb200_stem_conv3x3s2(x_nchw.data_ptr(), _TORCH2MAF[x_nchw.dtype], n, h, w, weight.data_ptr(),
                                       bias.data_ptr(), _act(act), dst.ref(), _stream()))


def dwconv(src: NHWC, weight: torch.Tensor, bias: torch.Tensor, k: int, act, dst: NHWC) -> None:
    check(lib().mafb200_dwconv(src.ref(), weight.data_ptr(), bias.data_ptr(), k, _act(act), dst.ref(), _stream()))


def dwconv_conv1x1(src: NHWC, dw_w: torch.Tensor, dw_b: torch.Tensor, k: int, act1, pw_w: torch.Tensor,
                   pw_b: torch.Tensor, act2, dst: NHWC) -> None:
    """Depth-wise k x k (+bias, act1) fused with the 1x1 conv (+bias, act2) that consumes it."""
    check(lib().mafb200_dwconv_conv1x1(src.ref(), dw_w.data_ptr(), dw_b.data_ptr(), k, _act(act1), pw_w.data_ptr(),
                                       pw_b.data_ptr(), _act(act2), dst.ref(), _stream()))


def maxpool2x2_conv1x1(src: NHWC, w: torch.Tensor, b: torch.Tensor, act, dst: NHWC) -> None:
    """2x2 stride-2 max pool fused with the 1x1 conv (+bias, act) that consumes it (MPRep's first branch)."""
    check(lib().mafb200_maxpool2x2_conv1x1(src.ref(), w.data_ptr(), b.data_ptr(), _act(act), dst.ref(), _stream()))


def maxpool2x2(src: NHWC, dst: NHWC) -> None:
    check(lib().mafb200_maxpool2x2(src.ref(), dst.ref(), _stream()))


def sppf_pool(src: NHWC, y1: NHWC, y2: NHWC, y3: NHWC) -> None:
    check(lib().mafb200_sppf_pool(src.ref(), y1.ref(), y2.ref(), y3.ref(), _stream()))


def upsample2x(src: NHWC, dst: NHWC) -> None:
    check(lib().mafb200_upsample2x(src.ref(), dst.ref(), _stream()))


def nchw_to_nhwc(x: torch.Tensor, dst: NHWC) -> None:
    assert x.is_contiguous() and x.is_cuda and tuple(x.shape) == (dst.n, dst.c, dst.h, dst.w)
    check(lib().mafb200_nchw_to_nhwc_f16(x.data_ptr(), _TORCH2MAF[x.dtype], dst.ref(), _stream()))


def nhwc_to_nchw(src: NHWC, out: torch.Tensor) -> None:
    assert out.is_contiguous() and out.is_cuda and tuple(out.shape) == (src.n, src.c, src.h, src.w)
    check(lib().mafb200_nhwc_f16_to_nchw(src.ref(), out.data_ptr(), _TORCH2MAF[out.dtype], _stream()))


def head_decode(cls_logits: Sequence[NHWC], reg: Sequence[NHWC], strides: Sequence[float], reg_max: int,
                pred: torch.Tensor, cls_is_prob: bool = False) -> None:
    nl = len(cls_logits)
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.is_cuda
    ca = (MafTensor * nl)(*[t.maf() for t in cls_logits])
    ra = (MafTensor * nl)(*[t.maf() for t in reg])
    st = (C.c_float * nl)(*[float(s) for s in strides])
    check(lib().mafb200_head_decode(ca, ra, st, nl, reg_max, int(cls_is_prob), pred.data_ptr(), _stream()))


def head_decode_detect(cls_logits: Sequence[NHWC], reg: Sequence[NHWC], strides: Sequence[float], reg_max: int,
                       boxes: torch.Tensor, conf_thres: float, multi_label: bool, class_filter: Optional[torch.Tensor],
                       workspace: torch.Tensor, pred: Optional[torch.Tensor] = None, cls_is_prob: bool = False) -> None:
    """Decode fused with the NMS threshold / compaction pass: boxes [B,A,4] + candidates in `workspace`."""
    nl = len(cls_logits)
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.is_cuda and boxes.shape[-1] == 4
    ca = (MafTensor * nl)(*[t.maf() for t in cls_logits])
    ra = (MafTensor * nl)(*[t.maf() for t in reg])
    st = (C.c_float * nl)(*[float(s) for s in strides])
    check(lib().mafb200_head_decode_detect(ca, ra, st, nl, reg_max, int(cls_is_prob),
                                           pred.data_ptr() if pred is not None else None, boxes.data_ptr(),
                                           float(conf_thres), int(bool(multi_label)),
                                           class_filter.data_ptr() if class_filter is not None else None,
                                           workspace.data_ptr(), workspace.numel() * workspace.element_size(), _stream()))


def nms_select(boxes: torch.Tensor, nc: int, iou_thres: float, agnostic: bool, max_det: int, max_nms: int,
               det: torch.Tensor, count: torch.Tensor, workspace: torch.Tensor) -> None:
    """Sort + greedy NMS over the candidates head_decode_detect left in `workspace`; boxes [B,A,4] or a pred tensor."""
    b, a, stride = boxes.shape
    check(lib().mafb200_nms_select(boxes.data_ptr(), stride, b, a, nc, float(iou_thres), int(bool(agnostic)), max_det,
                                   max_nms, det.data_ptr(), count.data_ptr(), workspace.data_ptr(),
                                   workspace.numel() * workspace.element_size(), _stream()))


def nms_select_packed(boxes: torch.Tensor, nc: int, iou_thres: float, agnostic: bool, max_det: int, max_nms: int,
                      packed: torch.Tensor, workspace: torch.Tensor) -> None:
    """nms_select writing rows of the all-gather layout: packed [B, row_floats] fp32 (dist.DetectionGather)."""
    b, a, stride = boxes.shape
    assert packed.dtype == torch.float32 and packed.dim() == 2 and packed.shape[0] == b and packed.stride(1) == 1
    check(lib().mafb200_nms_select_packed(boxes.data_ptr(), stride, b, a, nc, float(iou_thres), int(bool(agnostic)),
                                          max_det, max_nms, packed.data_ptr(), packed.stride(0), workspace.data_ptr(),
                                          workspace.numel() * workspace.element_size(), _stream()))


def nms_workspace_bytes(batch: int, anchors: int, nc: int) -> int:
    return int(lib().mafb200_nms_workspace_bytes(batch, anchors, nc))


def nms(pred: torch.Tensor, conf_thres: float, iou_thres: float, multi_label: bool, agnostic: bool,
        class_filter: Optional[torch.Tensor], max_det: int, max_nms: int, det: torch.Tensor, count: torch.Tensor,
        workspace: torch.Tensor) -> None:
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.is_cuda and pred.dim() == 3
    b, a, no = pred.shape
    check(lib().mafb200_nms(pred.data_ptr(), b, a, no - 5, float(conf_thres), float(iou_thres), int(bool(multi_label)),
                            int(bool(agnostic)), class_filter.data_ptr() if class_filter is not None else None,
                            max_det, max_nms, det.data_ptr(), count.data_ptr(), workspace.data_ptr(),
                            workspace.numel() * workspace.element_size(), _stream()))
