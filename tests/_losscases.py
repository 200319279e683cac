"""Seeded inputs of the loss / assigner tests (shared by the golden generator, the CPU pin and the GPU parity tests)."""
import torch

A = 8400
CASES = {
    # name: (batch, boxes per image, seed, prediction style)
    "sparse": (3, [5, 0, 12], 0, "random"),
    "crowded": (4, [30, 25, 1, 40], 1, "peaked"),
    "empty": (2, [0, 0], 2, "random"),
}


def make_case(name):
    """(pred_scores [B,A,80] probabilities, pred_distri [B,A,68] logits, targets [T,6]) on the CPU, fp32."""
    bs, ngt, seed, style = CASES[name]
    g = torch.Generator().manual_seed(seed)
    scores = torch.sigmoid(torch.randn(bs, A, 80, generator=g) * 1.5 - 3.0)
    distri = torch.randn(bs, A, 68, generator=g)
    if style == "peaked":  # sharper DFL distributions with larger expectations: bigger boxes, more anchors claimed twice
        distri = distri * 3.0 + torch.linspace(-1.0, 1.0, 17).repeat(4)
    rows = []
    for b in range(bs):
        for _ in range(ngt[b]):
            cx, cy = torch.rand(2, generator=g).tolist()
            w, h = (0.04 + 0.45 * torch.rand(2, generator=g)).tolist()
            rows.append([float(b), float(int(torch.randint(0, 80, (1,), generator=g))), cx, cy, w, h])
    targets = torch.tensor(rows, dtype=torch.float32).reshape(-1, 6)
    return scores, distri, targets
