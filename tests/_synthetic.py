"""Seeded synthetic inputs shared by the tests and tests/golden/make_golden.py."""
import torch


def synthetic_pred(b, a, nc, seed, dup=True):
    """A `[b, a, 5+nc]` prediction tensor with COCO-like candidate density (~0.1% of scores above
    0.03), random boxes, objectness 1, plus exact score ties / duplicate boxes."""
    g = torch.Generator().manual_seed(seed)
    pred = torch.zeros(b, a, 5 + nc)
    pred[..., 0:2] = torch.rand(b, a, 2, generator=g) * 640
    pred[..., 2:4] = torch.rand(b, a, 2, generator=g) * 200 + 4
    pred[..., 4] = 1.0
    pred[..., 5:] = torch.sigmoid(torch.randn(b, a, nc, generator=g) * 1.5 - 8)
    if dup and a > 120:
        pred[0, 100:110] = pred[0, 90:100]
        pred[0, 110:120, 5:] = pred[0, 90:100, 5:]
    return pred


def synthetic_image(b, h=640, w=640, seed=0, dtype=torch.float32):
    x = torch.rand(b, 3, h, w, generator=torch.Generator().manual_seed(seed))
    if dtype == torch.uint8:
        return (x * 255).round().to(torch.uint8)
    return x.to(dtype)
