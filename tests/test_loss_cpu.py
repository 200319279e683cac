"""Oracle of the training-side kernels (oracle/loss.py) against the UNMODIFIED reference and the golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as ol
from oracle import ref_loader
from tests import _losscases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _oracle_run(name, warmup=False):
    scores, distri, targets = _losscases.make_case(name)
    ps = scores.clone().requires_grad_()
    pd = distri.clone().requires_grad_()
    loss, items, asg = ol.compute_loss(ps, pd, targets, return_assignment=True, warmup=warmup)
    if torch.isfinite(loss):
        loss.backward()
        grads = (ps.grad, pd.grad)
    else:
        grads = (torch.zeros_like(ps), torch.zeros_like(pd))
    return loss.detach(), items, asg, grads


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("warmup", [False, True])
@pytest.mark.parametrize("name", list(_losscases.CASES))
def test_loss_oracle_bitexact_vs_reference(name, warmup):
    """Strongest pin: the restatement equals yolov6.models.loss.ComputeLoss bit for bit — loss, loss items, the
    assigner's four outputs and the autograd gradients w.r.t. both head outputs — with the formal (task-aligned) assigner
    and with the ATSS assigner of the warm-up epochs."""
    import sys

    sys.path.insert(0, GOLDEN)
    import make_golden_loss as mg

    cl = mg.reference_loss()
    scores, distri, targets = _losscases.make_case(name)
    loss_r, items_r, (t_labels, t_boxes, t_scores, fg), (gs_r, gd_r) = mg.run_reference(cl, scores, distri, targets, 0 if warmup else 5)
    loss_o, items_o, asg, (gs_o, gd_o) = _oracle_run(name, warmup)
    # float64 whenever there is a target (the numpy-built target tensor); fp32 on the reference's no-target early return
    assert loss_r.dtype == loss_o.dtype == (torch.float64 if targets.shape[0] else torch.float32)
    assert torch.equal(loss_r, loss_o) or (torch.isinf(loss_r) and torch.isinf(loss_o))
    assert torch.equal(items_r, items_o)
    assert torch.equal(fg.bool(), asg["fg_mask"])
    if fg.any():
        f = fg.bool()
        # the reference returns labels / boxes BEFORE loss.py:141,144 rewrite them; compare on the foreground anchors
        assert torch.equal(t_labels[f], asg["target_labels"][f])
        assert torch.equal(t_scores, asg["target_scores"])
        assert torch.equal(gs_r, gs_o) and torch.equal(gd_r, gd_o)


@pytest.mark.parametrize("name", list(_losscases.CASES) + ["sparse_atss", "crowded_atss"])
def test_loss_oracle_reproduces_golden(name):
    """The committed vectors (generated from the reference by tests/golden/make_golden_loss.py) from the seeds alone; this
    is the pin that travels to the GPU box.  Index sets exactly; floats to 1e-9 (another CPU's libm / vector width)."""
    gold = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    loss, items, asg, (gs, gd) = _oracle_run(name.replace("_atss", ""), warmup=name.endswith("_atss"))
    if np.isinf(gold["loss"]):
        assert torch.isinf(loss) and asg["fg_mask"].sum() == 0
        return
    np.testing.assert_allclose(loss.item(), gold["loss"], rtol=1e-9)
    np.testing.assert_allclose(items.numpy(), gold["loss_items"], rtol=1e-9)
    idx = torch.nonzero(asg["fg_mask"].reshape(-1)).squeeze(1)
    assert np.array_equal(idx.numpy(), gold["fg_index"])
    assert np.array_equal(asg["target_labels"].reshape(-1)[idx].numpy(), gold["fg_label"])
    pts, stride = ol.anchor_points()
    boxes_px = (asg["target_bboxes"] * stride).reshape(-1, 4)[idx]
    np.testing.assert_allclose(boxes_px.numpy(), gold["fg_box"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(asg["target_scores"].reshape(-1, 80)[idx].sum(-1).numpy(), gold["fg_score"], rtol=1e-9)
    np.testing.assert_allclose(gs.reshape(-1, 80)[idx].numpy(), gold["grad_scores_fg"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(gd.reshape(-1, 68)[idx].numpy(), gold["grad_distri_fg"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(gs.reshape(-1)[torch.from_numpy(gold["sample_index"])].numpy(), gold["grad_scores_sample"],
                               rtol=1e-5, atol=1e-12)


def test_preprocess_targets_layout():
    """loss.py:164-172: rows keep their order inside an image, padding rows are (-1, 0, 0, 0, 0), float64."""
    t = torch.tensor([[1, 3, 0.5, 0.5, 0.2, 0.4], [0, 7, 0.25, 0.75, 0.1, 0.1], [1, 5, 0.1, 0.2, 0.05, 0.3]])
    out = ol.preprocess_targets(t, 3)
    assert out.dtype == torch.float64 and out.shape == (3, 2, 5)
    assert out[1, 0, 0] == 3 and out[1, 1, 0] == 5 and out[0, 1, 0] == -1 and (out[2, :, 0] == -1).all()
    np.testing.assert_allclose(out[1, 0, 1:].numpy(), [(0.5 - 0.1) * 640, (0.5 - 0.2) * 640, (0.5 + 0.1) * 640, (0.5 + 0.2) * 640], rtol=1e-7)
    assert ol.preprocess_targets(torch.zeros(0, 6), 2).shape == (2, 0, 5)
