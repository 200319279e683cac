"""Imports the UNMODIFIED reference (yang-0201/MAF-YOLO) from /root/reference when it is present.

Used to (a) pin the oracle against the real reference in the build container and (b) generate the
golden vectors under tests/golden/ (tests/golden/make_golden.py).  /root/reference does not exist
on the GPU box: every caller must handle `available() == False`.

Three import-time dependencies of the reference are absent offline and are stubbed before import
(SURVEY.md appendix D): `timm.models.layers.DropPath` (yolov6/layers/common.py:1423),
`utils.general.LOGGER` and `utils.torch_utils.model_info` (yolov6/models/yolo.py:12-13).
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

REF_ROOT = os.environ.get("MAF_REFERENCE_ROOT", "/root/reference")
_state = {}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "yolov6"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with Model, common (layers module), fuse_model, non_max_suppression."""
    if _state:
        return _state["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.nn as nn

    sys.dont_write_bytecode = True  # the reference tree is read-only

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()

        def forward(self, x):
            return x

    if "timm" not in sys.modules:
        _stub("timm")
        _stub("timm.models")
        _stub("timm.models.layers", DropPath=DropPath)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    importlib.import_module("utils")  # the reference's light package __init__
    _stub("utils.general", LOGGER=logging.getLogger("maf_ref"))
    _stub("utils.torch_utils", model_info=lambda *a, **k: None)
    from yolov6.models.yolo import Model  # noqa: E402
    import yolov6.layers.common as common  # noqa: E402
    from yolov6.utils.torch_utils import fuse_model  # noqa: E402
    from yolov6.utils.nms import non_max_suppression  # noqa: E402

    ns = types.SimpleNamespace(Model=Model, common=common, fuse_model=fuse_model,
                               non_max_suppression=non_max_suppression, root=REF_ROOT)
    _state["ns"] = ns
    return ns


class _NS(dict):
    __getattr__ = dict.__getitem__


def build_model(variant: str = "n", num_classes: int = 80):
    """Reference `Model` in TRAIN form (eval mode), exactly as yolov6/models/yolo.py:127 builds it."""
    ns = load()
    cfg = _NS(model=_NS(build_type="yaml", yaml_file=os.path.join(ns.root, f"configs/yaml/MAF-YOLO-{variant}.yaml"),
                        head=_NS(num_layers=3, anchors=1, strides=[8, 16, 32], use_dfl=True, reg_max=16)))
    return ns.Model(cfg, channels=3, num_classes=num_classes, anchors=1).eval()


def to_deploy(model):
    """The reference's own deploy conversion (yolov6/core/evaler.py:93,101-109)."""
    ns = load()
    ns.fuse_model(model)
    for layer in model.modules():
        if isinstance(layer, ns.common.RepVGGBlock):
            layer.switch_to_deploy()
        if isinstance(layer, ns.common.UniRepLKNetBlock):
            layer.reparameterize()
    return model.eval()
