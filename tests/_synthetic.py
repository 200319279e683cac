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


def synthetic_scene(b, h=640, w=640, seed=7):
    """Images with structure at every anchor scale and no flat regions: piecewise-constant random grids of 64 / 32 /
    16 / 8 px cells (aligned with the stride-8/16/32 anchor cells, so neighbouring anchors see different content)
    plus pixel noise.  With signal-preserving weights (synth.random_state_dict(conv_gain=...)) every anchor then has
    its own features and the score maps have distinct peaks.  fp32 in [0, 1]."""
    import torch.nn.functional as F

    gen = torch.Generator().manual_seed(seed)
    x = torch.zeros(b, 3, h, w)
    for cell in (64, 32, 16, 8):
        r = torch.rand(b, 3, h // cell, w // cell, generator=gen) - 0.5
        x += 0.4 * F.interpolate(r, size=(h, w), mode="nearest")
    x += 0.06 * (torch.rand(b, 3, h, w, generator=gen) - 0.5)
    return (x + 0.5).clamp(0, 1)
