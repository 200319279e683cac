"""Load-time weight folding: reference `state_dict` (train OR deploy form) -> one (weight, bias) per
executed convolution, computed in float64 and rounded once when packed.

What the reference does at eval start-up (yolov6/core/evaler.py:91-113) as three in-place module
surgeries — `fuse_model` (Conv+BN, yolov6/utils/torch_utils.py:50-98), `RepVGGBlock.switch_to_deploy`
(3x3 + padded 1x1 branches, yolov6/layers/common.py:226-283) and `UniRepLKNetBlock.reparameterize`
(DilatedReparamBlock branches zero-padded into the k x k kernel, then the outer BN;
common.py:2940-2947,3033-3051,3085-3100) — is done here as pure tensor algebra keyed by the yaml
layer index.  BatchNorm eps is 1e-3 (torch_utils.py:43-45), not PyTorch's default.

Result keys (one per kernel launch that needs weights), i = yaml layer index:
  "{i}"                      RepVGGBlock 3x3           "{i}.block"            ConvWrapper 3x3
  "{i}.conv1", "{i}.conv2"   RepHDW / MPRep(1x1, RepVGG 3x3)
  "{i}.m.{j}.conv1|dw|one_conv"   DepthBottleneckUni
  "{i}.cv1", "{i}.cv2"       SPPF
  "{i}.stem|cls_dw|cls_s|cls_pred|reg_dw|reg_s|reg_pred"   Head_DepthUni
"""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

import torch
import torch.nn.functional as F

from .topology import Graph

DIL_BRANCHES = {9: [7, 5, 3], 7: [5, 3], 5: [3, 1], 3: [3, 1]}  # common.py:2985-3008
Folded = Dict[str, Tuple[torch.Tensor, torch.Tensor]]


class _Folder:
    def __init__(self, sd: Mapping[str, torch.Tensor], eps: float):
        self.sd = sd
        self.eps = eps

    def t(self, key: str) -> torch.Tensor:
        if key not in self.sd:
            raise KeyError(f"state_dict has no '{key}' (neither train-form nor deploy-form MAF-YOLO weights?)")
        return self.sd[key].detach().to("cpu", torch.float64)

    def has(self, key: str) -> bool:
        return key in self.sd

    def bn_affine(self, p: str):
        """BN(x) = x * scale + shift with running statistics."""
        scale = self.t(p + ".weight") / torch.sqrt(self.t(p + ".running_var") + self.eps)
        return scale, self.t(p + ".bias") - self.t(p + ".running_mean") * scale

    def conv_bn(self, conv_w: str, bn: str):
        w = self.t(conv_w)
        scale, shift = self.bn_affine(bn)
        return w * scale.view(-1, 1, 1, 1), shift

    def conv_module(self, p: str):
        """`Conv` (common.py:29-50): train form conv(no bias)+bn, deploy form conv with bias."""
        if self.has(p + ".bn.weight"):
            return self.conv_bn(p + ".conv.weight", p + ".bn")
        return self.t(p + ".conv.weight"), self.t(p + ".conv.bias")

    def repvgg(self, p: str):
        if self.has(p + ".rbr_reparam.weight"):
            return self.t(p + ".rbr_reparam.weight"), self.t(p + ".rbr_reparam.bias")
        k3, b3 = self.conv_bn(p + ".rbr_dense.conv.weight", p + ".rbr_dense.bn")
        k1, b1 = self.conv_bn(p + ".rbr_1x1.conv.weight", p + ".rbr_1x1.bn")
        if self.has(p + ".rbr_identity.weight"):
            raise NotImplementedError("RepVGG identity branch (stride 1) does not occur in MAF-YOLO")
        return k3 + F.pad(k1, [1, 1, 1, 1]), b3 + b1

    def unireplk(self, p: str, k: int):
        if not self.has(p + ".dwconv.origin_bn.weight"):  # already merged
            w, b = self.t(p + ".dwconv.lk_origin.weight"), self.t(p + ".dwconv.lk_origin.bias")
        else:
            w, b = self.conv_bn(p + ".dwconv.lk_origin.weight", p + ".dwconv.origin_bn")
            for kb in DIL_BRANCHES[k]:
                wb, bb = self.conv_bn(f"{p}.dwconv.dil_conv_k{kb}_1.weight", f"{p}.dwconv.dil_bn_k{kb}_1")
                pad = k // 2 - kb // 2
                w = w + F.pad(wb, [pad] * 4)
                b = b + bb
        if self.has(p + ".norm.weight"):  # outer BN not folded yet
            scale, shift = self.bn_affine(p + ".norm")
            w = w * scale.view(-1, 1, 1, 1)
            b = b * scale + shift
        return w, b


def fold_state_dict(graph: Graph, state_dict: Mapping[str, torch.Tensor], bn_eps: float = 1e-3) -> Folded:
    f = _Folder(state_dict, bn_eps)
    out: Folded = {}
    for l in graph.layers:
        p, i = f"backbone.{l.i}", str(l.i)
        if l.kind == "repvgg":
            out[i] = f.repvgg(p)
        elif l.kind == "rephdw":
            out[i + ".conv1"] = f.conv_module(p + ".conv1")
            for j in range(l.depth):
                q = f"{p}.m.{j}"
                out[f"{i}.m.{j}.conv1"] = f.conv_module(q + ".conv1")
                out[f"{i}.m.{j}.dw"] = f.unireplk(q + ".conv2", l.k)
                out[f"{i}.m.{j}.one_conv"] = f.conv_module(q + ".one_conv")
            out[i + ".conv2"] = f.conv_module(p + ".conv2")
        elif l.kind == "mprep":
            out[i + ".conv1"] = f.conv_module(p + ".conv1")
            out[i + ".conv2"] = f.repvgg(p + ".conv2")
        elif l.kind == "sppf":
            out[i + ".cv1"] = f.conv_module(p + ".cv1")
            out[i + ".cv2"] = f.conv_module(p + ".cv2")
        elif l.kind == "convw":
            out[i + ".block"] = f.conv_module(p + ".block")
        elif l.kind == "head":
            out[i + ".stem"] = f.conv_module(p + ".stem")
            for br in ("cls", "reg"):
                out[f"{i}.{br}_dw"] = f.unireplk(f"{p}.{br}_conv", l.k)
                out[f"{i}.{br}_s"] = f.conv_module(f"{p}.{br}_conv_s")
                out[f"{i}.{br}_pred"] = (f.t(f"{p}.{br}_pred.weight"), f.t(f"{p}.{br}_pred.bias"))
    return out


def deploy_param_count(folded: Folded) -> int:
    """Parameters of the deploy-form network (README.md:24-26 reports 3.76 M / 8.55 M / 23.7 M)."""
    return sum(w.numel() + b.numel() for w, b in folded.values())
