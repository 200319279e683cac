"""Seeded synthetic weights in the reference's TRAIN-form `state_dict` schema.

There is no network for released checkpoints, so benchmarks and tests run random-init weights of
the exact architecture: convs use PyTorch's default Conv2d init (kaiming_uniform, a=sqrt(5)), and —
unlike a fresh reference model — BatchNorm statistics and the head prediction weights are
randomised, otherwise BN folding would be untested (fresh BN is the identity) and every class
score would be the constant 0.01 (the reference zero-initialises cls_pred/reg_pred weights,
yolov6/layers/common.py:1313-1322).

The result loads into the reference model with `load_state_dict(strict=True)`.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .topology import Graph

DIL_BRANCHES = {9: [7, 5, 3], 7: [5, 3], 5: [3, 1], 3: [3, 1]}  # yolov6/layers/common.py:2985-3008


def random_state_dict(graph: Graph, seed: int = 0, head_std: float = 0.35, cls_bias: float = -4.595,
                      conv_gain: float = 1.0, reg_peak: float = None, reg_sharp: float = 1.0,
                      reg_std: float = None) -> Dict[str, torch.Tensor]:
    """conv_gain scales every conv's init bound.  1.0 = PyTorch's default init, under which a signal decays by ~0.4x
    per layer, so after 30+ layers the heads see almost only the last layers' biases (predictions barely depend on
    the image).  sqrt(6) ~ 2.45 gives He-normalised weights (std sqrt(2 / fan_in)): activations keep their scale
    through the whole network, as in a trained model — the setting of the conditioned parity fixtures
    (tests/golden/make_golden_cond.py), where every layer's error reaches the output.
    reg_peak / reg_sharp / reg_std shape the DFL head: reg_pred.bias of bin i becomes -reg_sharp * (i - reg_peak)^2
    (the softmax expectation of every box side then sits near reg_peak grid units, i.e. boxes of ~2 * reg_peak cells
    instead of the ~17-cell boxes a constant bias gives) and reg_pred.weight ~ N(0, reg_std)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k, groups=1):
        fan_in = (ci // groups) * k * k
        bound = conv_gain / math.sqrt(fan_in)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        sd[name] = (torch.rand(co, ci // groups, k, k, generator=g) * 2 - 1) * bound

    def bn(name, c):
        sd[name + ".weight"] = torch.rand(c, generator=g) + 0.5
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) + 0.5
        sd[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    def conv_mod(p, ci, co, k=1):
        conv(p + ".conv.weight", co, ci, k)
        bn(p + ".bn", co)

    def repvgg(p, ci, co):
        conv(p + ".rbr_dense.conv.weight", co, ci, 3)
        bn(p + ".rbr_dense.bn", co)
        conv(p + ".rbr_1x1.conv.weight", co, ci, 1)
        bn(p + ".rbr_1x1.bn", co)

    def unireplk(p, c, k):
        conv(p + ".dwconv.lk_origin.weight", c, c, k, groups=c)
        bn(p + ".dwconv.origin_bn", c)
        for kb in DIL_BRANCHES[k]:
            conv(f"{p}.dwconv.dil_conv_k{kb}_1.weight", c, c, kb, groups=c)
            bn(f"{p}.dwconv.dil_bn_k{kb}_1", c)
        bn(p + ".norm", c)

    for l in graph.layers:
        p = f"backbone.{l.i}"
        if l.kind == "repvgg":
            repvgg(p, l.c_in[0], l.c_out)
        elif l.kind == "rephdw":
            c_ = l.c_hidden
            conv_mod(p + ".conv1", l.c_in[0], 2 * c_)
            for j in range(l.depth):
                q = f"{p}.m.{j}"
                conv_mod(q + ".conv1", c_, l.expand * c_)
                unireplk(q + ".conv2", l.expand * c_, l.k)
                conv_mod(q + ".one_conv", l.expand * c_, c_)
            conv_mod(p + ".conv2", (2 + l.depth) * c_, l.c_out)
        elif l.kind == "mprep":
            conv_mod(p + ".conv1", l.c_in[0], l.c_out // 2)
            repvgg(p + ".conv2", l.c_in[0], l.c_out // 2)
        elif l.kind == "sppf":
            conv_mod(p + ".cv1", l.c_in[0], l.c_hidden)
            conv_mod(p + ".cv2", 4 * l.c_hidden, l.c_out)
        elif l.kind == "convw":
            conv_mod(p + ".block", l.c_in[0], l.c_out, 3)
        elif l.kind == "head":
            c = l.c_out
            conv_mod(p + ".stem", l.c_in[0], c)
            unireplk(p + ".cls_conv", c, l.k)
            conv_mod(p + ".cls_conv_s", c, c)
            unireplk(p + ".reg_conv", c, l.k)
            conv_mod(p + ".reg_conv_s", c, c)
            sd[p + ".cls_pred.weight"] = torch.randn(graph.nc, c, 1, 1, generator=g) * head_std
            sd[p + ".cls_pred.bias"] = torch.full((graph.nc,), cls_bias)
            sd[p + ".reg_pred.weight"] = torch.randn(4 * (l.reg_max + 1), c, 1, 1, generator=g) * (
                head_std if reg_std is None else reg_std)
            if reg_peak is None:
                sd[p + ".reg_pred.bias"] = torch.full((4 * (l.reg_max + 1),), 1.0)
            else:
                bins = torch.arange(l.reg_max + 1, dtype=torch.float32)
                sd[p + ".reg_pred.bias"] = (-reg_sharp * (bins - reg_peak) ** 2).repeat(4)
    reg_max = graph.layers[graph.head_layers[0]].reg_max
    sd["detect.proj"] = torch.linspace(0, reg_max, reg_max + 1)
    sd["detect.proj_conv.weight"] = sd["detect.proj"].view(1, reg_max + 1, 1, 1).clone()
    return sd
