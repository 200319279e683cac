#!/usr/bin/env python3
"""Generates tests/golden/preprocess.npz from the UNMODIFIED reference + its cv2 (run in the build container):
letterbox (yolov6/data/data_augment.py:53-83) and Inferer.precess_image (yolov6/core/inferer.py:168-178).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pre.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from tests._precases import CASES, image  # noqa: E402
import make_golden_post as mg  # noqa: E402


def main():
    Evaler, Inferer = mg.load_reference()
    from yolov6.data.data_augment import letterbox

    out = {}
    for i, (h, w, ns, auto, scaleup) in enumerate(CASES):
        im = image(i, h, w)
        lb, r, (dw, dh) = letterbox(im, ns, auto=auto, scaleup=scaleup, stride=32)
        if i < 4:
            out[f"lb_{i}"] = lb  # full arrays for a few cases, digests for all (keeps the fixture small)
        out[f"shape_{i}"] = np.array(lb.shape, dtype=np.int64)
        out[f"sha_{i}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(lb).tobytes()).digest(), dtype=np.uint8)
        out[f"meta_{i}"] = np.array([r, dw, dh], dtype=np.float64)
    img, _ = Inferer.precess_image(image(0, 97, 131), 160, 32, False)
    out["precess_0"] = img.numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)
    print("wrote preprocess.npz", len(out))


if __name__ == "__main__":
    main()
