"""GPU parity of mafb200_scale_detections (through maf_yolo_b200.postprocess) with the oracle and the golden
vectors from the unmodified reference: bit-exact boxes, identical COCO json records."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "postprocess.npz")


@pytest.mark.parametrize("name", ["eval", "exact"])
def test_coco_records_and_boxes(cuda_device, name):
    from maf_yolo_b200 import postprocess as post
    from oracle import postprocess as opost
    from tests._postcases import make_cases, pad_batch

    gold = np.load(GOLD)
    c = make_cases()[name]
    det, cnt = pad_batch(c["outputs"])
    det, cnt = det.to(cuda_device), cnt.to(cuda_device)
    got = post.convert_to_coco_format(det, cnt, c["img_shape"], c["paths"], c["shapes"], c["ids"], True, c["scale_exact"])
    assert got == json.loads(bytes(gold[name + "_coco_json"]).decode())
    assert got == opost.convert_to_coco_format([d.clone() for d in c["outputs"]], c["img_shape"], c["paths"], c["shapes"],
                                               c["ids"], True, c["scale_exact"])
    xyxy = post.scale_coords(c["img_shape"], det, cnt, c["shapes"], c["scale_exact"]).cpu().numpy()
    resc = post.rescale(c["img_shape"], det, cnt, [s[0] for s in c["shapes"]]).cpu().numpy()
    for i, d in enumerate(c["outputs"]):
        n = len(d)
        if n:
            assert np.array_equal(xyxy[i, :n, :4], gold[f"{name}_xyxy_{i}"]), f"image {i}: scale_coords differs"
            assert np.array_equal(resc[i, :n, :4], gold[f"{name}_rescale_{i}"]), f"image {i}: rescale differs"
            assert np.array_equal(xyxy[i, :n, 4:], d[:, 4:].numpy())
        assert (xyxy[i, n:] == 0).all(), "padding rows must be zero"


def test_on_real_nms_output_and_recip_form(cuda_device):
    """End of the path: forward -> NMS -> device rescale == oracle on the same NMS output; the torch-CUDA
    (reciprocal-multiply) form differs from the CPU form by at most 1 ulp."""
    import maf_yolo_b200 as mb
    from maf_yolo_b200 import postprocess as post, synth, topology
    from oracle import postprocess as opost
    from tests._postcases import COCO91, _letterbox_meta
    from tests._synthetic import synthetic_image

    g = topology.build_graph("n")
    model = mb.from_state_dict(synth.random_state_dict(g, seed=0), "n")
    x = synthetic_image(2, seed=0).to(cuda_device)
    det, cnt = mb.non_max_suppression_padded(model(x)[0], 0.03, 0.65, multi_label=True)
    shapes = [_letterbox_meta(427, 640), _letterbox_meta(640, 480)]
    paths = ["000000000139.jpg", "000000000285.jpg"]
    got = post.convert_to_coco_format(det, cnt, (640, 640), paths, shapes, COCO91)
    outs = [det[i, :int(cnt[i])].cpu().clone() for i in range(2)]
    assert got == opost.convert_to_coco_format(outs, (640, 640), paths, shapes, COCO91) and len(got) > 0
    a = post.scale_coords((640, 640), det, cnt, shapes).cpu()
    b = post.scale_coords((640, 640), det, cnt, shapes, recip_mul=True).cpu()
    assert (a - b).abs().max().item() <= 1e-4 * a.abs().max().item()
