"""Oracle of the post-NMS step (oracle/postprocess.py) against golden vectors generated from the unmodified
reference (tests/golden/make_golden_post.py) and, when /root/reference is present, against the reference's own
Evaler / Inferer methods directly.  Bit-exact (fp32 index/box arithmetic, json text)."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import postprocess as opost
from oracle import ref_loader
from tests._postcases import make_cases

GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess.npz")


@pytest.mark.parametrize("name", ["eval", "exact"])
def test_oracle_matches_golden(name):
    gold = np.load(GOLD)
    c = make_cases()[name]
    outputs = [d.clone() for d in c["outputs"]]
    res = opost.convert_to_coco_format(outputs, c["img_shape"], c["paths"], c["shapes"], c["ids"], True, c["scale_exact"])
    want = json.loads(bytes(gold[name + "_coco_json"]).decode())
    assert res == want
    for i, d in enumerate(c["outputs"]):
        if len(d) == 0:
            continue
        xy = opost.scale_coords(c["img_shape"], d[:, :4].clone(), c["shapes"][i][0], c["shapes"][i][1], c["scale_exact"])
        assert np.array_equal(xy.numpy(), gold[f"{name}_xyxy_{i}"])
        rs = opost.rescale(c["img_shape"], d[:, :4].clone(), c["shapes"][i][0])
        assert np.array_equal(rs.numpy(), gold[f"{name}_rescale_{i}"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_matches_reference_methods():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_post as mg

    Evaler, Inferer = mg.load_reference()
    c = make_cases()["eval"]
    fake = types.SimpleNamespace(scale_exact=False, is_coco=True, ids=c["ids"])
    fake.scale_coords = types.MethodType(Evaler.scale_coords, fake)
    fake.box_convert = types.MethodType(Evaler.box_convert, fake)
    imgs = torch.zeros((len(c["outputs"]), 3, 640, 640))
    want = Evaler.convert_to_coco_format(fake, [d.clone() for d in c["outputs"]], imgs, c["paths"], c["shapes"], c["ids"])
    got = opost.convert_to_coco_format([d.clone() for d in c["outputs"]], (640, 640), c["paths"], c["shapes"], c["ids"])
    assert got == want and len(got) == sum(len(d) for d in c["outputs"])
    d = c["outputs"][2]
    assert torch.equal(Inferer.rescale((640, 640), d[:, :4].clone(), (1080, 1920)),
                       opost.rescale((640, 640), d[:, :4].clone(), (1080, 1920)))
