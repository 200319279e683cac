"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the reference's MAF-YOLO
forward -> decode path, in plain torch functional ops on NCHW tensors.

Two executable forms, both driven by a reference-format `state_dict`:

  forward_train_form(spec, sd, x)   the reference `Model.forward` in eval() mode on the TRAIN-form
                                    module tree (multi-branch RepVGG / DilatedReparam, BatchNorm with
                                    running statistics) — yolov6/models/yolo.py:179-209 and the layer
                                    classes cited per function below.  This is the parity oracle: it
                                    involves no weight folding at all.
  fold_deploy(spec, sd)             the reference's deploy conversion (fuse_model +
  forward_deploy(spec, dd, x)       switch_to_deploy + reparameterize, yolov6/core/evaler.py:93-109)
                                    restated, and the deploy-form forward.  Used as the timed CPU
                                    baseline ("port") and to cross-check the product's own fold.

Pinned (tests/test_oracle_cpu.py, run where /root/reference exists) against the real
reference modules loaded with the same state_dict, and against committed golden vectors
(tests/golden/, made by tests/golden/make_golden.py from the reference).  The reference ships no
tests / golden vectors of its own, so beyond that parity is unpinned.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # initialize_weights sets eps on every BatchNorm2d: yolov6/utils/torch_utils.py:43-45
StateDict = Dict[str, torch.Tensor]

# kernel-size -> small-kernel branches of DilatedReparamBlock in this fork (all dilation 1),
# yolov6/layers/common.py:2985-3008
DIL_BRANCHES = {9: [7, 5, 3], 7: [5, 3], 5: [3, 1], 3: [3, 1]}


# ------------------------------------------------------------------------------------------------
# topology: rows in the reference yaml schema + parse_model's channel rules (yolo.py:15-120)
# ------------------------------------------------------------------------------------------------
def variant_rows(variant: str) -> dict:
    """The rows of configs/yaml/MAF-YOLO-{n,s,m}.yaml, regenerated from their differences."""
    width = {"n": 0.375, "s": 0.5, "m": 0.75}[variant]
    bb_n = {"n": [1, 1, 1, 1], "s": [2, 2, 2, 2], "m": [2, 4, 4, 2]}[variant]
    bb_c = {"n": [48, 96, 192, 384], "s": [64, 128, 256, 512], "m": [96, 192, 384, 768]}[variant]
    nd = {"n": 1, "s": 2, "m": 3}[variant]
    cw = {"n": [96, 64, 64, 128, 128], "s": [128, 96, 96, 192, 192], "m": [256, 192, 192, 192, 192]}[variant]
    hd = {"n": [192, 128, 128, 128, 128, 192], "s": [256, 192, 192, 192, 192, 256],
          "m": [512, 384, 384, 256, 384, 384]}[variant]
    heads = {"n": [341, 341, 512], "s": [384, 384, 512], "m": [341, 512, 512]}[variant]
    U = [-1, 1, "nn.Upsample", [None, 2, "nearest"]]
    backbone = [[-1, 1, "RepVGGBlock", [64, 3, 2]], [-1, 1, "RepVGGBlock", [128, 3, 2]],
                [-1, bb_n[0], "RepHDW", [bb_c[0], True, 0.5, 3, 3]], [-1, 1, "MPRep", [256]],
                [-1, bb_n[1], "RepHDW", [bb_c[1], True, 0.5, 5, 3]], [-1, 1, "MPRep", [512]],
                [-1, bb_n[2], "RepHDW", [bb_c[2], True, 0.5, 7, 3]], [-1, 1, "MPRep", [1024]],
                [-1, bb_n[3], "RepHDW", [bb_c[3], True, 0.5, 9, 3]], [-1, 1, "SPPF", [1024, 5]]]
    neck = [[6, 1, "ConvWrapper", [cw[0], 3, 2]], [[-1, 9], 1, "Concat", [1]],
            [-1, nd, "RepHDW", [hd[0], False, 0.5, 9, 3]], U,
            [4, 1, "ConvWrapper", [cw[1], 3, 2]], [[-1, 6, -2], 1, "Concat", [1]],
            [-1, nd, "RepHDW", [hd[1], False, 0.5, 7, 3]], U,
            [2, 1, "ConvWrapper", [cw[2], 3, 2]], [[-1, 4, -2], 1, "Concat", [1]],
            [-1, nd, "RepHDW", [hd[2], False, 0.5, 5, 3]],
            [[-1, 17], 1, "Concat", [1]], [-1, nd, "RepHDW", [hd[3], False, 0.5, 5, 3]],
            [-1, 1, "ConvWrapper", [cw[3], 3, 2]], [20, 1, "ConvWrapper", [cw[3], 3, 2]],
            [[-2, -1, 16, 13], 1, "Concat", [1]], [-1, nd, "RepHDW", [hd[4], False, 0.5, 7, 3]],
            [-1, 1, "ConvWrapper", [cw[4], 3, 2]], [16, 1, "ConvWrapper", [cw[4], 3, 2]],
            [[-2, -1, 12], 1, "Concat", [1]], [-1, nd, "RepHDW", [hd[5], False, 0.5, 9, 3]]]
    head = [[22, 1, "Head_DepthUni", [heads[0], 16, 5]], [26, 1, "Head_DepthUni", [heads[1], 16, 7]],
            [30, 1, "Head_DepthUni", [heads[2], 16, 9]], [[31, 32, 33], 1, "Out", []]]
    return dict(depth_multiple=1, width_multiple=width, backbone=backbone, neck=neck, effidehead=head)


def _make_divisible(x, d):
    return math.ceil(x / d) * d  # yolo.py:220-222


def parse_model(rows: dict, nc: int = 80, ch: int = 3) -> List[dict]:
    """yolo.py:15-120 for the module types MAF-YOLO uses; returns one dict per yaml row."""
    gd, gw = rows["depth_multiple"], rows["width_multiple"]
    chs: List[int] = []
    spec = []
    for i, (f, n, m, args) in enumerate(rows["backbone"] + rows["neck"] + rows["effidehead"]):
        n = max(round(n * gd), 1) if n > 1 else n  # yolo.py:27
        src = [f] if isinstance(f, int) else list(f)
        c_in = [ch if (i == 0 and s == -1) else chs[s] for s in src]  # negative = relative (python indexing)
        m = m.split(".")[-1]
        d = dict(i=i, f=f, type=m)
        if m == "RepVGGBlock":  # yolo.py:28-32
            d.update(c1=c_in[0], c2=_make_divisible(args[0] * gw, 4))
        elif m == "SPPF":
            d.update(c1=c_in[0], c2=_make_divisible(args[0] * gw, 4))
        elif m == "RepHDW":  # yolo.py:36-40: args -> [c1, c2, depth, shortcut, expansion, kersize, depth_expansion]
            c2 = args[0]
            d.update(c1=c_in[0], c2=c2, depth=n, c_=int(c2 * args[2]), k=args[3], expand=args[4])
        elif m == "Concat":  # yolo.py:43-44
            d.update(c2=sum(c_in))
        elif m == "Head_DepthUni":  # yolo.py:56-59
            d.update(c1=c_in[0], c2=_make_divisible(args[0] * gw, 8), reg_max=args[1], k=args[2], nc=nc)
        elif m == "ConvWrapper":  # yolo.py:64-67
            d.update(c1=c_in[0], c2=args[0])
        elif m == "MPRep":  # yolo.py:92-96
            d.update(c1=c_in[0], c2=_make_divisible(args[0] * gw, 8))
        elif m == "Upsample":
            d.update(c2=c_in[0])
        elif m == "Out":
            d.update(c2=0)
        else:
            raise NotImplementedError(m)
        spec.append(d)
        chs.append(d["c2"])
    return spec


def savelist(spec: Sequence[dict]) -> List[int]:
    save = []
    for d in spec:
        f = d["f"]
        save.extend(x % d["i"] for x in ([f] if isinstance(f, int) else f) if x != -1)  # yolo.py:115
    return sorted(save)


# ------------------------------------------------------------------------------------------------
# train-form modules in eval mode
# ------------------------------------------------------------------------------------------------
def _bn(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False,
                        0.0, BN_EPS)


def _conv_mod(sd, p, x, stride=1):
    """`Conv`: conv -> bn -> SiLU (common.py:29-50); padding = k // 2."""
    w = sd[p + ".conv.weight"]
    return F.silu(_bn(sd, p + ".bn", F.conv2d(x, w, None, stride, w.shape[-1] // 2)))


def _repvgg(sd, p, x):
    """RepVGGBlock train form, stride 2, no identity branch (common.py:214-224)."""
    d = _bn(sd, p + ".rbr_dense.bn", F.conv2d(x, sd[p + ".rbr_dense.conv.weight"], None, 2, 1))
    o = _bn(sd, p + ".rbr_1x1.bn", F.conv2d(x, sd[p + ".rbr_1x1.conv.weight"], None, 2, 0))
    return F.relu(d + o)


def _unireplk(sd, p, x, k):
    """UniRepLKNetBlock train form: norm(DilatedReparamBlock(x)) (common.py:3018-3031,3079-3083)."""
    c = x.shape[1]
    out = _bn(sd, p + ".dwconv.origin_bn", F.conv2d(x, sd[p + ".dwconv.lk_origin.weight"], None, 1, k // 2, 1, c))
    for kb in DIL_BRANCHES[k]:
        w = sd[f"{p}.dwconv.dil_conv_k{kb}_1.weight"]
        out = out + _bn(sd, f"{p}.dwconv.dil_bn_k{kb}_1", F.conv2d(x, w, None, 1, kb // 2, 1, c))
    return _bn(sd, p + ".norm", out)


def _rephdw(sd, p, x, d):
    """RepHDW.forward (common.py:938-946) with DepthBottleneckUni.forward (common.py:920-927)."""
    y = _conv_mod(sd, p + ".conv1", x)
    outs = list(y.split((d["c_"], d["c_"]), 1))
    for j in range(d["depth"]):
        q = f"{p}.m.{j}"
        t = _conv_mod(sd, q + ".conv1", outs[-1])
        t = F.silu(_unireplk(sd, q + ".conv2", t, d["k"]))
        outs.append(_conv_mod(sd, q + ".one_conv", t))
    return _conv_mod(sd, p + ".conv2", torch.cat(outs, 1))


def _mprep(sd, p, x):
    """MPRep.forward (common.py:787-792)."""
    return torch.cat([_conv_mod(sd, p + ".conv1", F.max_pool2d(x, 2, 2)), _repvgg(sd, p + ".conv2", x)], 1)


def _sppf(sd, p, x):
    """SPPF.forward (common.py:123-129)."""
    x = _conv_mod(sd, p + ".cv1", x)
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    return _conv_mod(sd, p + ".cv2", torch.cat((x, y1, y2, F.max_pool2d(y2, 5, 1, 2)), 1))


def _head(sd, p, x, d):
    """Head_DepthUni.forward (common.py:1325-1336): returns (stem, sigmoid(cls), reg)."""
    x = _conv_mod(sd, p + ".stem", x)
    cf = _conv_mod(sd, p + ".cls_conv_s", _unireplk(sd, p + ".cls_conv", x, d["k"]))
    cls = torch.sigmoid(F.conv2d(cf, sd[p + ".cls_pred.weight"], sd[p + ".cls_pred.bias"]))
    rf = _conv_mod(sd, p + ".reg_conv_s", _unireplk(sd, p + ".reg_conv", x, d["k"]))
    reg = F.conv2d(rf, sd[p + ".reg_pred.weight"], sd[p + ".reg_pred.bias"])
    return x, cls, reg


# ------------------------------------------------------------------------------------------------
# decode (Detect_yaml eval branch)
# ------------------------------------------------------------------------------------------------
def generate_anchors_eval(sizes: Sequence[Tuple[int, int]], strides: Sequence[float], offset: float = 0.5):
    """anchor_generator.py:11-25 (is_eval=True)."""
    pts, sts = [], []
    for (h, w), s in zip(sizes, strides):
        sx = torch.arange(end=w) + offset
        sy = torch.arange(end=h) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack([xx, yy], -1).to(torch.float).reshape(-1, 2))
        sts.append(torch.full((h * w, 1), s, dtype=torch.float))
    return torch.cat(pts), torch.cat(sts)


def dist2bbox_xywh(distance, anchor_points):
    """general.py:29-40 with box_format='xywh'."""
    lt, rb = torch.split(distance, 2, -1)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    return torch.cat([(x1y1 + x2y2) / 2, x2y2 - x1y1], -1)


def detect_eval(head_outs, strides=(8, 16, 32), reg_max=16, nc=80):
    """Detect_yaml.forward eval branch (yolo.py:355-396).  head_outs: list of (stem, cls_prob, reg)."""
    anchor_points, stride_tensor = generate_anchors_eval([o[0].shape[2:] for o in head_outs], strides)
    # yolo.py:328-330; a parameter of the model, so `model.half()` (evaler.py:112) casts it with the convs while the
    # anchor points stay fp32 (anchor_generator.py:18) and promote the boxes back to fp32
    proj = torch.linspace(0, reg_max, reg_max + 1).view(1, reg_max + 1, 1, 1).to(head_outs[0][2])
    cls_l, reg_l = [], []
    for stem, cls, reg in head_outs:
        b, _, h, w = stem.shape
        l = h * w
        r = reg.reshape([-1, 4, reg_max + 1, l]).permute(0, 2, 1, 3)
        r = F.conv2d(F.softmax(r, dim=1), proj)  # yolo.py:377-378
        cls_l.append(cls.reshape([b, nc, l]))
        reg_l.append(r.reshape([b, 4, l]))
    cls_all = torch.cat(cls_l, -1).permute(0, 2, 1)
    reg_all = torch.cat(reg_l, -1).permute(0, 2, 1)
    boxes = dist2bbox_xywh(reg_all, anchor_points)
    boxes = boxes * stride_tensor
    return torch.cat([boxes, torch.ones((boxes.shape[0], boxes.shape[1], 1), dtype=boxes.dtype, device=boxes.device), cls_all], -1)


# ------------------------------------------------------------------------------------------------
# whole model
# ------------------------------------------------------------------------------------------------
def _run_graph(spec, x, layer_fn, keep: Optional[Sequence[int]] = None):
    """Model.forward's executor loop (yolo.py:189-200).  Returns (final x, {i: tensor} for `keep`)."""
    save = set(savelist(spec))
    y: List[Optional[torch.Tensor]] = []
    kept = {}
    for d in spec:
        f = d["f"]
        if f != -1:
            x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
        x = layer_fn(d, x)
        y.append(x if d["i"] in save else None)
        if keep is not None and d["i"] in keep:
            kept[d["i"]] = x
    return x, kept


def forward_train_form(spec, sd: StateDict, x: torch.Tensor, keep=None, decode=True):
    """x [B,3,H,W] fp32 in [0,1] -> pred [B, A, 5+nc] (eval output of the reference Model)."""

    def layer(d, x):
        p, t = f"backbone.{d['i']}", d["type"]
        if t == "RepVGGBlock":
            return _repvgg(sd, p, x)
        if t == "RepHDW":
            return _rephdw(sd, p, x, d)
        if t == "MPRep":
            return _mprep(sd, p, x)
        if t == "SPPF":
            return _sppf(sd, p, x)
        if t == "ConvWrapper":
            return _conv_mod(sd, p + ".block", x, 2)
        if t == "Concat":
            return torch.cat(x, 1)
        if t == "Upsample":
            return F.interpolate(x, scale_factor=2.0, mode="nearest")
        if t == "Head_DepthUni":
            return _head(sd, p, x, d)
        if t == "Out":
            return list(x)
        raise NotImplementedError(t)

    with torch.no_grad():
        outs, kept = _run_graph(spec, x, layer, keep)
        heads = [d for d in spec if d["type"] == "Head_DepthUni"]
        pred = detect_eval(outs, reg_max=heads[0]["reg_max"], nc=heads[0]["nc"]) if decode else outs
    return (pred, kept) if keep is not None else pred


# ---- deploy form -----------------------------------------------------------------------------------
def _fuse(w, sd, bn, b=None):
    """fuse_conv_and_bn / fuse_bn (torch_utils.py:50-82, common.py:2636-2645)."""
    std = (sd[bn + ".running_var"] + BN_EPS).sqrt()
    t = sd[bn + ".weight"] / std
    bias = sd[bn + ".bias"] - sd[bn + ".running_mean"] * t
    if b is not None:
        bias = bias + b * t
    return w * t.reshape(-1, 1, 1, 1), bias


def _fold_conv(sd, p):
    return _fuse(sd[p + ".conv.weight"], sd, p + ".bn")


def _fold_repvgg(sd, p):
    """get_equivalent_kernel_bias (common.py:226-264), no identity branch."""
    k3, b3 = _fuse(sd[p + ".rbr_dense.conv.weight"], sd, p + ".rbr_dense.bn")
    k1, b1 = _fuse(sd[p + ".rbr_1x1.conv.weight"], sd, p + ".rbr_1x1.bn")
    return k3 + F.pad(k1, [1, 1, 1, 1]), b3 + b1


def _fold_unireplk(sd, p, k):
    """merge_dilated_branches (common.py:3033-3051) then fold the outer norm (common.py:3085-3100)."""
    wk, bk = _fuse(sd[p + ".dwconv.lk_origin.weight"], sd, p + ".dwconv.origin_bn")
    for kb in DIL_BRANCHES[k]:
        w, b = _fuse(sd[f"{p}.dwconv.dil_conv_k{kb}_1.weight"], sd, f"{p}.dwconv.dil_bn_k{kb}_1")
        pad = k // 2 - kb // 2
        wk = wk + F.pad(w, [pad] * 4)
        bk = bk + b
    std = (sd[p + ".norm.running_var"] + BN_EPS).sqrt()
    g = sd[p + ".norm.weight"] / std
    return wk * g.view(-1, 1, 1, 1), sd[p + ".norm.bias"] + (bk - sd[p + ".norm.running_mean"]) * g


def fold_deploy(spec, sd: StateDict, dtype=torch.float32) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """Deploy-form (weight, bias) per conv, keyed by the reference module path."""
    sd = {k: v.to(dtype) for k, v in sd.items() if v.is_floating_point()}
    dd = {}
    for d in spec:
        p, t = f"backbone.{d['i']}", d["type"]
        if t == "RepVGGBlock":
            dd[p] = _fold_repvgg(sd, p)
        elif t == "RepHDW":
            dd[p + ".conv1"] = _fold_conv(sd, p + ".conv1")
            for j in range(d["depth"]):
                q = f"{p}.m.{j}"
                dd[q + ".conv1"] = _fold_conv(sd, q + ".conv1")
                dd[q + ".conv2"] = _fold_unireplk(sd, q + ".conv2", d["k"])
                dd[q + ".one_conv"] = _fold_conv(sd, q + ".one_conv")
            dd[p + ".conv2"] = _fold_conv(sd, p + ".conv2")
        elif t == "MPRep":
            dd[p + ".conv1"] = _fold_conv(sd, p + ".conv1")
            dd[p + ".conv2"] = _fold_repvgg(sd, p + ".conv2")
        elif t == "SPPF":
            dd[p + ".cv1"] = _fold_conv(sd, p + ".cv1")
            dd[p + ".cv2"] = _fold_conv(sd, p + ".cv2")
        elif t == "ConvWrapper":
            dd[p + ".block"] = _fold_conv(sd, p + ".block")
        elif t == "Head_DepthUni":
            dd[p + ".stem"] = _fold_conv(sd, p + ".stem")
            for br in ("cls", "reg"):
                dd[f"{p}.{br}_conv"] = _fold_unireplk(sd, f"{p}.{br}_conv", d["k"])
                dd[f"{p}.{br}_conv_s"] = _fold_conv(sd, f"{p}.{br}_conv_s")
                dd[f"{p}.{br}_pred"] = (sd[f"{p}.{br}_pred.weight"], sd[f"{p}.{br}_pred.bias"])
    return dd


def forward_deploy(spec, dd, x: torch.Tensor, keep=None, decode=True):
    """Deploy-form forward: every conv carries its bias, activation follows directly."""

    def cv(p, x, act, stride=1, groups=1):
        w, b = dd[p]
        y = F.conv2d(x, w, b, stride, w.shape[-1] // 2, 1, groups)
        return F.silu(y) if act == "silu" else F.relu(y) if act == "relu" else y

    def layer(d, x):
        p, t = f"backbone.{d['i']}", d["type"]
        if t == "RepVGGBlock":
            return cv(p, x, "relu", 2)
        if t == "RepHDW":
            y = cv(p + ".conv1", x, "silu")
            outs = list(y.split((d["c_"], d["c_"]), 1))
            for j in range(d["depth"]):
                q = f"{p}.m.{j}"
                tt = cv(q + ".conv1", outs[-1], "silu")
                tt = cv(q + ".conv2", tt, "silu", groups=tt.shape[1])
                outs.append(cv(q + ".one_conv", tt, "silu"))
            return cv(p + ".conv2", torch.cat(outs, 1), "silu")
        if t == "MPRep":
            return torch.cat([cv(p + ".conv1", F.max_pool2d(x, 2, 2), "silu"), cv(p + ".conv2", x, "relu", 2)], 1)
        if t == "SPPF":
            x = cv(p + ".cv1", x, "silu")
            y1 = F.max_pool2d(x, 5, 1, 2)
            y2 = F.max_pool2d(y1, 5, 1, 2)
            return cv(p + ".cv2", torch.cat((x, y1, y2, F.max_pool2d(y2, 5, 1, 2)), 1), "silu")
        if t == "ConvWrapper":
            return cv(p + ".block", x, "silu", 2)
        if t == "Concat":
            return torch.cat(x, 1)
        if t == "Upsample":
            return F.interpolate(x, scale_factor=2.0, mode="nearest")
        if t == "Head_DepthUni":
            x = cv(p + ".stem", x, "silu")
            cf = cv(p + ".cls_conv_s", cv(p + ".cls_conv", x, None, groups=x.shape[1]), "silu")
            rf = cv(p + ".reg_conv_s", cv(p + ".reg_conv", x, None, groups=x.shape[1]), "silu")
            return x, torch.sigmoid(cv(p + ".cls_pred", cf, None)), cv(p + ".reg_pred", rf, None)
        if t == "Out":
            return list(x)
        raise NotImplementedError(t)

    with torch.no_grad():
        outs, kept = _run_graph(spec, x, layer, keep)
        heads = [d for d in spec if d["type"] == "Head_DepthUni"]
        pred = detect_eval(outs, reg_max=heads[0]["reg_max"], nc=heads[0]["nc"]) if decode else outs
    return (pred, kept) if keep is not None else pred
