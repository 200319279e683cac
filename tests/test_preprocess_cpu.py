"""Oracle of the pre-processing step (oracle/preprocess.py, pure numpy) against golden vectors generated from the
unmodified reference + cv2 (tests/golden/make_golden_pre.py) and, in the build container, against cv2.resize and
the reference's letterbox directly.  Bit-exact (uint8)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import preprocess as opre
from oracle import ref_loader
from tests._precases import CASES, image

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz")


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_oracle_matches_golden():
    gold = np.load(GOLD)
    for i, (h, w, ns, auto, scaleup) in enumerate(CASES):
        lb, r, (dw, dh) = opre.letterbox(image(i, h, w), ns, auto=auto, scaleup=scaleup, stride=32)
        assert tuple(lb.shape) == tuple(gold[f"shape_{i}"]), i
        assert np.array_equal(_sha(lb), gold[f"sha_{i}"]), f"case {i}: letterbox differs from the reference"
        if f"lb_{i}" in gold:
            assert np.array_equal(lb, gold[f"lb_{i}"])
        assert np.array_equal(np.array([r, dw, dh], dtype=np.float64), gold[f"meta_{i}"])
    chw, flt = opre.precess_image(image(0, 97, 131), 160, 32, False)
    assert np.array_equal(flt, gold["precess_0"])


def test_resize_matches_cv2_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for (h0, w0, dw, dh) in [(480, 640, 640, 480), (1080, 1920, 640, 360), (375, 500, 640, 480), (281, 500, 640, 360),
                             (720, 1280, 640, 360), (200, 300, 640, 427), (31, 17, 160, 96), (500, 333, 107, 160)]:
        img = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
        assert np.array_equal(opre.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_matches_reference_letterbox():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_post as mg

    mg.load_reference()
    from yolov6.data.data_augment import letterbox

    for i, (h, w, ns, auto, scaleup) in enumerate(CASES[:6]):
        im = image(i, h, w)
        want = letterbox(im, ns, auto=auto, scaleup=scaleup, stride=32, return_int=True)
        got = opre.letterbox(im, ns, auto=auto, scaleup=scaleup, stride=32, return_int=True)
        assert np.array_equal(want[0], got[0]) and want[1] == got[1] and tuple(want[2]) == tuple(got[2])
