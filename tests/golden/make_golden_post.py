#!/usr/bin/env python3
"""Generates tests/golden/postprocess.npz from the UNMODIFIED reference (run in the build container):
Evaler.scale_coords / box_convert / convert_to_coco_format (yolov6/core/evaler.py:382-442) and
Inferer.rescale (yolov6/core/inferer.py:181-195) on seeded synthetic detections.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_post.py
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests._postcases import make_cases  # noqa: E402


def load_reference():
    ref_loader.load()

    def stub(name, **a):
        m = types.ModuleType(name)
        m.__dict__.update(a)
        sys.modules[name] = m

    stub("pycocotools"); stub("pycocotools.coco", COCO=object); stub("pycocotools.cocoeval", COCOeval=object)
    stub("yolov6.data.data_load", create_dataloader=None)
    stub("albumentations")
    from yolov6.core.evaler import Evaler
    from yolov6.core.inferer import Inferer
    return Evaler, Inferer


def main():
    Evaler, Inferer = load_reference()
    cases = make_cases()
    out = {}
    for name, c in cases.items():
        fake = types.SimpleNamespace(scale_exact=c["scale_exact"], is_coco=True, ids=c["ids"])
        fake.scale_coords = types.MethodType(Evaler.scale_coords, fake)
        fake.box_convert = types.MethodType(Evaler.box_convert, fake)
        outputs = [d.clone() for d in c["outputs"]]
        imgs = torch.zeros((len(outputs), 3) + tuple(c["img_shape"]))
        res = Evaler.convert_to_coco_format(fake, outputs, imgs, c["paths"], c["shapes"], c["ids"])
        out[name + "_coco_json"] = np.frombuffer(json.dumps(res).encode(), dtype=np.uint8)
        # scale_coords alone (xyxy), per image
        for i, d in enumerate(c["outputs"]):
            if len(d):
                xy = d[:, :4].clone()
                Evaler.scale_coords(fake, c["img_shape"], xy, c["shapes"][i][0], c["shapes"][i][1])
                out[f"{name}_xyxy_{i}"] = xy.numpy()
                rs = d[:, :4].clone()
                Inferer.rescale(c["img_shape"], rs, c["shapes"][i][0])
                out[f"{name}_rescale_{i}"] = rs.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "postprocess.npz"), **out)
    print("wrote postprocess.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
