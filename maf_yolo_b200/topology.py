"""MAF-YOLO topology: the three variants (N/S/M) and the reference's yaml row format, resolved to a
static list of `Layer`s with every channel count known.

The reference builds its graph by eval()-ing module names from a yaml
(yolov6/models/yolo.py:15-120, configs/yaml/MAF-YOLO-{n,s,m}.yaml).  Here the same information is
data: `variant_rows()` generates the `[from, number, module, args]` rows of a variant from a small
parameter table, `rows_from_yaml()` accepts a reference yaml (file or dict) for custom models, and
`resolve()` applies the reference's channel rules:

  RepVGGBlock / SPPF : c2 = make_divisible(args[0] * width, 4)          yolo.py:28-32
  RepHDW             : c2 = args[0] (NOT width-scaled), number -> depth  yolo.py:36-40
  ConvWrapper        : c2 = args[0] (not scaled)                         yolo.py:64-67
  MPRep / Head       : c2 = make_divisible(args[0] * width, 8)           yolo.py:56-59,92-96
  Concat             : c2 = sum of inputs                                yolo.py:43-44
  number             : max(round(n * depth), 1) if n > 1 else n          yolo.py:27
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

# (width_multiple, backbone RepHDW (number, channels) x4, neck RepHDW number,
#  neck channels: cw10, L12, cw14, L16, cw18, L20, L22, cw23/24, L26, cw27/28, L30, head widths)
_VARIANTS = {
    "n": dict(width=0.375, bb=[(1, 48), (1, 96), (1, 192), (1, 384)], nd=1,
              neck=dict(cw10=96, l12=192, cw14=64, l16=128, cw18=64, l20=128, l22=128, cw23=128, l26=128, cw27=128,
                        l30=192), heads=[341, 341, 512]),
    "s": dict(width=0.5, bb=[(2, 64), (2, 128), (2, 256), (2, 512)], nd=2,
              neck=dict(cw10=128, l12=256, cw14=96, l16=192, cw18=96, l20=192, l22=192, cw23=192, l26=192, cw27=192,
                        l30=256), heads=[384, 384, 512]),
    "m": dict(width=0.75, bb=[(2, 96), (4, 192), (4, 384), (2, 768)], nd=3,
              neck=dict(cw10=256, l12=512, cw14=192, l16=384, cw18=192, l20=384, l22=256, cw23=192, l26=384, cw27=192,
                        l30=384), heads=[341, 512, 512]),
}
VARIANTS = tuple(_VARIANTS)


def make_divisible(x: float, divisor: int) -> int:
    return math.ceil(x / divisor) * divisor  # yolo.py:220-222


def variant_rows(variant: str) -> dict:
    """Rows of MAF-YOLO-{n,s,m} in the reference's yaml schema (backbone 10, neck 21, head 3 + Out)."""
    v = _VARIANTS[variant.lower()]
    nk, nd = v["neck"], v["nd"]
    ks = [3, 5, 7, 9]
    rows = [[-1, 1, "RepVGGBlock", [64, 3, 2]], [-1, 1, "RepVGGBlock", [128, 3, 2]]]
    mp = [256, 512, 1024]
    for stage, (num, ch) in enumerate(v["bb"]):
        rows.append([-1, num, "RepHDW", [ch, True, 0.5, ks[stage], 3]])
        if stage < 3:
            rows.append([-1, 1, "MPRep", [mp[stage]]])
    rows.append([-1, 1, "SPPF", [1024, 5]])  # 9

    def hdw(frm, ch, k):
        return [frm, nd, "RepHDW", [ch, False, 0.5, k, 3]]

    up = [-1, 1, "nn.Upsample", [None, 2, "nearest"]]
    neck = [
        [6, 1, "ConvWrapper", [nk["cw10"], 3, 2]],      # 10
        [[-1, 9], 1, "Concat", [1]],                    # 11
        hdw(-1, nk["l12"], 9),                          # 12
        list(up),                                       # 13
        [4, 1, "ConvWrapper", [nk["cw14"], 3, 2]],      # 14
        [[-1, 6, -2], 1, "Concat", [1]],                # 15
        hdw(-1, nk["l16"], 7),                          # 16
        list(up),                                       # 17
        [2, 1, "ConvWrapper", [nk["cw18"], 3, 2]],      # 18
        [[-1, 4, -2], 1, "Concat", [1]],                # 19
        hdw(-1, nk["l20"], 5),                          # 20
        [[-1, 17], 1, "Concat", [1]],                   # 21
        hdw(-1, nk["l22"], 5),                          # 22  P3 out
        [-1, 1, "ConvWrapper", [nk["cw23"], 3, 2]],     # 23
        [20, 1, "ConvWrapper", [nk["cw23"], 3, 2]],     # 24
        [[-2, -1, 16, 13], 1, "Concat", [1]],           # 25
        hdw(-1, nk["l26"], 7),                          # 26  P4 out
        [-1, 1, "ConvWrapper", [nk["cw27"], 3, 2]],     # 27
        [16, 1, "ConvWrapper", [nk["cw27"], 3, 2]],     # 28
        [[-2, -1, 12], 1, "Concat", [1]],               # 29
        hdw(-1, nk["l30"], 9),                          # 30  P5 out
    ]
    head = [[22, 1, "Head_DepthUni", [v["heads"][0], 16, 5]], [26, 1, "Head_DepthUni", [v["heads"][1], 16, 7]],
            [30, 1, "Head_DepthUni", [v["heads"][2], 16, 9]], [[31, 32, 33], 1, "Out", []]]
    return dict(depth_multiple=1, width_multiple=v["width"], backbone=rows, neck=neck, effidehead=head)


def rows_from_yaml(src: Union[str, dict]) -> dict:
    """Accepts a reference topology yaml (path or already-loaded dict, as `Model.yaml` holds it)."""
    if isinstance(src, dict):
        return src
    import yaml

    with open(src, encoding="ascii", errors="ignore") as f:
        return yaml.safe_load(f)


@dataclass
class Layer:
    i: int                      # yaml row index == index in the reference's `Model.backbone` Sequential
    kind: str                   # repvgg | rephdw | mprep | sppf | convw | concat | upsample | head | out
    frm: List[int]              # absolute source layer indices (-1 = network input for layer 0)
    c_in: List[int]             # channels of each source
    c_out: int
    stride_total: int           # input-image pixels per output pixel (8/16/32 at the heads)
    depth: int = 1              # RepHDW: number of DepthBottleneckUni
    k: int = 0                  # depth-wise kernel size (RepHDW / head)
    c_hidden: int = 0           # RepHDW: c_ ; SPPF: c_ ; head: width
    expand: int = 3             # RepHDW depth_expansion (mid = expand * c_)
    reg_max: int = 16
    extra: dict = field(default_factory=dict)


@dataclass
class Graph:
    layers: List[Layer]
    nc: int
    save: List[int]             # the reference's savelist (yolo.py:115)
    head_layers: List[int]
    strides: List[int]
    variant: Optional[str] = None


def resolve(rows: dict, nc: int = 80, ch: int = 3, variant: Optional[str] = None) -> Graph:
    gd, gw = rows["depth_multiple"], rows["width_multiple"]
    spec = list(rows["backbone"]) + list(rows["neck"]) + list(rows["effidehead"])
    chs: List[int] = []
    strides: List[int] = []
    layers: List[Layer] = []
    save: List[int] = []
    for i, (f, n, m, args) in enumerate(spec):
        m = m if isinstance(m, str) else getattr(m, "__name__", str(m))
        args = list(args)
        n = max(round(n * gd), 1) if n > 1 else n
        f_list = [f] if isinstance(f, int) else list(f)
        frm = [(i - 1 if x == -1 else (x if x >= 0 else i + x)) for x in f_list]
        c_in = [ch if s < 0 else chs[s] for s in frm]
        s_in = [1 if s < 0 else strides[s] for s in frm]
        save.extend(x % i for x in f_list if x != -1)
        name = m.split(".")[-1]
        if name == "RepVGGBlock":
            assert args[1] == 3 and args[2] == 2, "only 3x3 stride-2 RepVGG blocks occur in MAF-YOLO"
            lay = Layer(i, "repvgg", frm, c_in, make_divisible(args[0] * gw, 4), s_in[0] * 2)
        elif name == "RepHDW":
            c2 = args[0]
            expansion, kers, dexp = args[2], args[3], args[4]  # [c2, shortcut, expansion, kersize, depth_expansion]
            c_ = int(c2 * expansion)
            lay = Layer(i, "rephdw", frm, c_in, c2, s_in[0], depth=n, k=kers, c_hidden=c_, expand=dexp)
        elif name == "MPRep":
            lay = Layer(i, "mprep", frm, c_in, make_divisible(args[0] * gw, 8), s_in[0] * 2)
        elif name == "SPPF":
            assert args[1] == 5
            lay = Layer(i, "sppf", frm, c_in, make_divisible(args[0] * gw, 4), s_in[0], c_hidden=c_in[0] // 2)
        elif name == "ConvWrapper":
            assert args[1] == 3 and args[2] == 2, "only 3x3 stride-2 ConvWrapper occurs in MAF-YOLO"
            lay = Layer(i, "convw", frm, c_in, args[0], s_in[0] * 2)
        elif name == "Concat":
            assert len(set(s_in)) == 1
            lay = Layer(i, "concat", frm, c_in, sum(c_in), s_in[0])
        elif name == "Upsample":
            assert args[1] == 2 and args[2] == "nearest"
            assert s_in[0] % 2 == 0
            lay = Layer(i, "upsample", frm, c_in, c_in[0], s_in[0] // 2)
        elif name == "Head_DepthUni":
            c2 = make_divisible(args[0] * gw, 8)
            lay = Layer(i, "head", frm, c_in, c2, s_in[0], k=args[2], c_hidden=c2, reg_max=args[1])
        elif name == "Out":
            lay = Layer(i, "out", frm, c_in, 0, 0)
        else:
            raise NotImplementedError(f"module '{m}' (row {i}) is not part of the MAF-YOLO hot path")
        layers.append(lay)
        chs.append(lay.c_out)
        strides.append(lay.stride_total)
    heads = [l.i for l in layers if l.kind == "head"]
    return Graph(layers, nc, sorted(save), heads, [layers[h].stride_total for h in heads], variant)


def build_graph(variant_or_yaml: Union[str, dict] = "n", nc: int = 80) -> Graph:
    if isinstance(variant_or_yaml, str) and variant_or_yaml.lower() in _VARIANTS:
        return resolve(variant_rows(variant_or_yaml), nc, variant=variant_or_yaml.lower())
    return resolve(rows_from_yaml(variant_or_yaml), nc)
