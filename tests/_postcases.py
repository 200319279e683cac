"""Seeded synthetic post-NMS cases shared by the golden generator, the oracle tests and the GPU tests."""
import numpy as np
import torch

COCO91 = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32, 33, 34,
          35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64,
          65, 67, 70, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90]  # evaler.py coco80_to_coco91_class


def _letterbox_meta(h0, w0, new=(640, 640), scaleup=False):
    """ratio / pad exactly as yolov6/data/data_augment.py:53-76 + datasets.py:211-215 compute them (python floats)."""
    # load_image resizes the long side to img_size first (datasets.py:277-300): r0 = 640 / max(h0, w0)
    r0 = new[0] / max(h0, w0)
    h, w = (int(h0 * r0), int(w0 * r0)) if r0 != 1 else (h0, w0)
    r = min(new[0] / h, new[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(w * r)), int(round(h * r))
    dw, dh = (new[1] - new_unpad[0]) / 2, (new[0] - new_unpad[1]) / 2
    return (h0, w0), ((h * r / h0, w * r / w0), (dw, dh))


def make_cases():
    g = torch.Generator().manual_seed(123)
    sizes = [(480, 640), (427, 640), (640, 480), (375, 500), (333, 500), (1080, 1920), (612, 612), (500, 281)]
    cases = {}
    for name, scale_exact in (("eval", False), ("exact", True)):
        outputs, shapes, paths = [], [], []
        for i, (h0, w0) in enumerate(sizes):
            n = [37, 0, 300, 5, 1, 112, 64, 9][i]
            cxcy = torch.rand(n, 2, generator=g) * 700 - 30          # some boxes partly outside the 640 canvas
            wh = torch.rand(n, 2, generator=g) * 300 + 1
            box = torch.cat([cxcy - wh / 2, cxcy + wh / 2], 1)
            score = torch.rand(n, 1, generator=g)
            cls = torch.randint(0, 80, (n, 1), generator=g).float()
            outputs.append(torch.cat([box, score, cls], 1).float())
            shapes.append(_letterbox_meta(h0, w0))
            paths.append(f"/data/coco/images/val2017/{100000 + 37 * i:012d}.jpg")
        cases[name] = dict(outputs=outputs, shapes=shapes, paths=paths, ids=COCO91, img_shape=(640, 640),
                           scale_exact=scale_exact)
    return cases


def pad_batch(outputs, max_det=300):
    b = len(outputs)
    det = torch.zeros((b, max_det, 6), dtype=torch.float32)
    cnt = torch.zeros((b,), dtype=torch.int32)
    for i, d in enumerate(outputs):
        det[i, :len(d)] = d
        cnt[i] = len(d)
    return det, cnt
