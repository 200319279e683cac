"""GPU parity of mafb200_letterbox_u8 (through maf_yolo_b200.preprocess) with the oracle and the golden vectors
from the unmodified reference: bit-exact uint8 images, identical ratio / padding."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def test_letterbox_cases(cuda_device):
    from maf_yolo_b200 import preprocess as pre
    from oracle import preprocess as opre
    from tests._precases import CASES, image

    gold = np.load(GOLD)
    for i, (h, w, ns, auto, scaleup) in enumerate(CASES):
        im = image(i, h, w)
        chw, r, (dw, dh) = pre.letterbox(im, ns, auto=auto, scaleup=scaleup, stride=32, device=cuda_device)
        got = chw.permute(1, 2, 0).contiguous().cpu().numpy()
        want = opre.letterbox(im, ns, auto=auto, scaleup=scaleup, stride=32)
        assert np.array_equal(got, want[0]), f"case {i}: differs from the oracle"
        sha = np.frombuffer(hashlib.sha256(got.tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(sha, gold[f"sha_{i}"]), f"case {i}: differs from the reference"
        assert np.array_equal(np.array([r, dw, dh], dtype=np.float64), gold[f"meta_{i}"])


def test_precess_image_and_full_size(cuda_device):
    from maf_yolo_b200 import preprocess as pre
    from oracle import preprocess as opre
    from tests._precases import image

    gold = np.load(GOLD)
    img, src = pre.precess_image(image(0, 97, 131), 160, 32, False, device=cuda_device, as_uint8=False)
    # the float form divides on the device, where torch computes x * (1/255) instead of x / 255: <= 1 ulp apart
    assert torch.allclose(img.cpu(), torch.from_numpy(gold["precess_0"]), rtol=2e-7, atol=0)
    u8, _ = pre.precess_image(image(0, 97, 131), 160, 32, device=cuda_device)
    assert u8.dtype == torch.uint8 and torch.equal(u8.float().cpu() / 255, torch.from_numpy(gold["precess_0"]))
    # COCO-sized sources at the real network size, incl. up-scaling and exact-fit (no resize) cases
    rng = np.random.default_rng(11)
    for (h0, w0) in [(480, 640), (427, 640), (1080, 1920), (375, 500), (640, 640), (333, 500)]:
        im = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
        got = pre.letterbox(im, 640, auto=False, scaleup=True, device=cuda_device)[0].permute(1, 2, 0).cpu().numpy()
        assert np.array_equal(got, opre.letterbox(im, 640, auto=False, scaleup=True)[0]), (h0, w0)


def test_batch_feeds_the_model(cuda_device):
    """letterbox_batch -> uint8 NCHW RGB batch -> model(uint8) == model(float(oracle preprocessing))."""
    import maf_yolo_b200 as mb
    from maf_yolo_b200 import preprocess as pre, synth, topology
    from oracle import preprocess as opre

    rng = np.random.default_rng(5)
    ims = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in [(480, 640), (500, 375)]]
    batch, shapes = pre.letterbox_batch(ims, 640, device=cuda_device)
    assert batch.shape == (2, 3, 640, 640) and batch.dtype == torch.uint8
    want = np.stack([np.ascontiguousarray(opre.letterbox(im, 640, auto=False, scaleup=False)[0].transpose(2, 0, 1)[::-1])
                     for im in ims])
    assert np.array_equal(batch.cpu().numpy(), want)
    assert shapes[0][0] == (480, 640) and shapes[0][1][1] == (0.0, 80.0)
    # the dataloader's `shapes` when the images are load_image outputs of larger files (datasets.py:215, 280-301)
    orig = [(960, 1280), (1000, 750)]
    sizes = [pre.load_image_size(h0, w0, 640) for h0, w0 in orig]
    assert sizes == [(480, 640), (640, 480)]
    ims2 = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (h, w) in sizes]
    _, shapes2 = pre.letterbox_batch(ims2, 640, device=cuda_device, orig_shapes=orig)
    for (h, w), (h0, w0), sh in zip(sizes, orig, shapes2):
        _, ratio, pad = opre.letterbox(np.zeros((h, w, 3), np.uint8), 640, auto=False, scaleup=False)
        assert sh == ((h0, w0), ((h * ratio / h0, w * ratio / w0), pad))
    g = topology.build_graph("n")
    model = mb.from_state_dict(synth.random_state_dict(g, seed=0), "n", use_cuda_graph=False)
    a = model(batch)[0].clone()
    b = model(torch.from_numpy(want).to(cuda_device).float() / 255)[0]
    assert (a - b).abs().max().item() <= 2e-2  # same kernel; uint8/255 vs pre-divided fp32 differ by fp32 rounding only
