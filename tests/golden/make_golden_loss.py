"""Golden vectors of the training-side kernels, generated from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_loss.py
Runs `yolov6.models.loss.ComputeLoss` (formal assigner: epoch_num >= warmup_epoch) on the CPU — the reference hard-codes
`.cuda()` on two parameter-free sub-modules (loss.py:52-53), which is neutralised by making `nn.Module.cuda` the identity
for the construction — on the seeded cases of tests/_losscases.py, with the formal assigner (epoch 5: loss_<case>.npz) and, for
the two cases with targets, with the ATSS warm-up assigner (epoch 0: loss_<case>_atss.npz), and stores per case:
  loss, loss_items                        float64, exactly as returned
  fg_index [F]                            flat (image * 8400 + anchor) indices of the foreground anchors
  fg_label [F], fg_box [F,4], fg_score [F] assigned class / box (pixels) / normalised target score of those anchors
  grad_scores_fg [F,80], grad_distri_fg [F,68]   autograd gradients at the foreground anchors
  grad_scores_sample [S], sample_index [S]       and at S = 4096 seeded flat positions of pred_scores
  grad_scores_sum, grad_distri_abs_sum           float64 checksums over the whole gradient tensors
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402
from tests import _losscases  # noqa: E402


def reference_loss():
    ref_loader.load()
    cuda = torch.nn.Module.cuda
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        from yolov6.models.loss import ComputeLoss
        return ComputeLoss()
    finally:
        torch.nn.Module.cuda = cuda


def run_reference(cl, scores, distri, targets, epoch=5):
    """-> loss, items, assigner outputs, (grad_scores, grad_distri).  epoch 5: formal (task-aligned) assigner; epoch 0:
    ATSS warm-up assigner (loss.py:83)."""
    ps = scores.clone().requires_grad_()
    pd = distri.clone().requires_grad_()
    bs = ps.shape[0]
    feats = [torch.zeros(bs, 1, 80, 80), torch.zeros(bs, 1, 40, 40), torch.zeros(bs, 1, 20, 20)]
    captured = {}
    assigner = cl.formal_assigner if epoch >= cl.warmup_epoch else cl.warmup_assigner
    orig = assigner.forward

    def spy(*a, **k):
        out = orig(*a, **k)
        captured["out"] = [o.clone() for o in out]
        return out

    assigner.forward = spy
    try:
        loss, items = cl((feats, ps, pd), targets.clone(), epoch, 1)
    finally:
        assigner.forward = orig
    if torch.isfinite(loss):
        loss.backward()
        grads = (ps.grad.clone(), pd.grad.clone())
    else:
        grads = (torch.zeros_like(ps), torch.zeros_like(pd))
    return loss.detach(), items, captured["out"], grads


def main():
    cl = reference_loss()
    for name, epoch, suffix in [(n, 5, "") for n in _losscases.CASES] + [(n, 0, "_atss") for n in ("sparse", "crowded")]:
        scores, distri, targets = _losscases.make_case(name)
        loss, items, (t_labels, t_boxes, t_scores, fg), (gs, gd) = run_reference(cl, scores, distri, targets, epoch)
        fg_flat = fg.reshape(-1).bool()
        idx = torch.nonzero(fg_flat).squeeze(1)
        samp = torch.randint(0, gs.numel(), (4096,), generator=torch.Generator().manual_seed(7))
        np.savez_compressed(
            os.path.join(HERE, f"loss_{name}{suffix}.npz"),
            loss=np.float64(loss.item()), loss_items=items.double().numpy(),
            fg_index=idx.numpy().astype(np.int64),
            fg_label=t_labels.reshape(-1)[idx].numpy().astype(np.int64),
            fg_box=t_boxes.reshape(-1, 4)[idx].double().numpy(),
            fg_score=t_scores.reshape(-1, t_scores.shape[-1])[idx].sum(-1).double().numpy(),
            grad_scores_fg=gs.reshape(-1, gs.shape[-1])[idx].numpy(), grad_distri_fg=gd.reshape(-1, gd.shape[-1])[idx].numpy(),
            sample_index=samp.numpy(), grad_scores_sample=gs.reshape(-1)[samp].numpy(),
            grad_scores_sum=np.float64(gs.double().sum().item()), grad_distri_abs_sum=np.float64(gd.double().abs().sum().item()))
        print(name + suffix, "loss", loss.item(), "items", items.tolist(), "fg", int(idx.numel()),
              "multi-claimed anchors would show as fg with IoU-resolved boxes; T =", targets.shape[0])


if __name__ == "__main__":
    main()
