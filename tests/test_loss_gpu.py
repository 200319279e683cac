"""Training-side CUDA kernels (csrc/loss.cu through maf_yolo_b200.loss.ComputeLoss -> mafb200_detect_loss) against the
oracle (oracle/loss.py, pinned bit-exactly to the reference) and the reference-generated golden vectors.

Index work (which anchors are foreground, which box each one gets) must be EXACT; float64 target scores 1e-9; the loss
1e-6 (fp32 log / log1p / exp of the CUDA math library vs the host's); gradients 2e-4 relative to the largest entry."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as ol
from tests import _losscases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run_cuda(dev, name, override, gt_cap=None, grads=True, warmup=False):
    from maf_yolo_b200.loss import ComputeLoss

    scores, distri, targets = _losscases.make_case(name)
    ps = scores.to(dev).requires_grad_(grads)
    pd = distri.to(dev).requires_grad_(grads)
    crit = ComputeLoss(warmup_epoch=3 if warmup else 0)  # epoch 0 below: ATSS while warming up (loss.py:83)
    if override:
        pts, stride = ol.anchor_points()
        crit._boxes_override = ol.decode_boxes(distri, pts / stride).contiguous().to(dev)
    loss, items = crit((None, ps, pd), targets.to(dev), 0, 0, gt_cap=gt_cap)
    if grads and torch.isfinite(loss):
        loss.backward()
    torch.cuda.synchronize()
    return loss, items, crit.last, ps.grad, pd.grad, targets


def _check_against_golden(name, loss, items, last, gs, gd, targets, loss_rtol, exact_boxes=True, score_rtol=None):
    gold = np.load(os.path.join(GOLDEN, f"loss_{name}.npz"))
    fg = last["fg_mask"].reshape(-1).cpu()
    idx = torch.nonzero(fg).squeeze(1)
    assert np.array_equal(idx.numpy(), gold["fg_index"]), (
        f"{name}: foreground set differs: {len(idx)} vs {len(gold['fg_index'])} anchors, "
        f"{len(np.setxor1d(idx.numpy(), gold['fg_index']))} in the symmetric difference")
    bs = last["fg_mask"].shape[0]
    gt = ol.preprocess_targets(targets, bs)
    img = idx // 8400
    gi = last["target_gt_idx"].reshape(-1).cpu()[idx].long()
    assert np.array_equal(gt[img, gi, 0].long().numpy(), gold["fg_label"])
    assert np.array_equal(gt[img, gi, 1:].numpy(), gold["fg_box"])  # float64 boxes of the assigned targets, bit for bit
    np.testing.assert_allclose(last["target_score"].reshape(-1).cpu()[idx].numpy(), gold["fg_score"], rtol=score_rtol or (1e-9 if exact_boxes else 1e-4))
    np.testing.assert_allclose(loss.item(), gold["loss"], rtol=loss_rtol)
    np.testing.assert_allclose(items.cpu().numpy(), gold["loss_items"], rtol=loss_rtol)
    s = last["scalars"].cpu()
    assert int(s[5]) == len(gold["fg_index"]) and int(s[6]) == 0
    if gs is not None:
        gsf = gs.reshape(-1, 80).cpu()[idx].numpy()
        gdf = gd.reshape(-1, 68).cpu()[idx].numpy()
        for got, want, what in ((gsf, gold["grad_scores_fg"], "d/d pred_scores (fg rows)"), (gdf, gold["grad_distri_fg"], "d/d pred_distri (fg rows)"),
                                (gs.reshape(-1).cpu()[torch.from_numpy(gold["sample_index"])].numpy(), gold["grad_scores_sample"], "d/d pred_scores (sample)")):
            scale = np.abs(want).max()
            err = np.abs(got - want).max() / scale
            print(f"{name} {what}: max err / max |ref| = {err:.2e}")
            assert err < 2e-4, what
        np.testing.assert_allclose(gs.double().sum().item(), gold["grad_scores_sum"], rtol=1e-4)
        np.testing.assert_allclose(gd.double().abs().sum().item(), gold["grad_distri_abs_sum"], rtol=1e-4)
        bg = torch.nonzero(~fg.bool()).squeeze(1)[:2000]
        assert (gd.reshape(-1, 68).cpu()[bg] == 0).all()  # background anchors get no box gradient


@pytest.mark.parametrize("name", ["sparse", "crowded"])
def test_assign_and_loss_with_reference_boxes(cuda_device, name):
    """Assigner + loss fed the oracle's decoded boxes: removes the fp32 softmax ulps from the comparison."""
    loss, items, last, gs, gd, targets = _run_cuda(cuda_device, name, override=True)
    _check_against_golden(name, loss, items, last, gs, gd, targets, loss_rtol=1e-6)


@pytest.mark.parametrize("name", ["sparse", "crowded"])
def test_detect_loss_end_to_end(cuda_device, name):
    """The public call: own DFL decode -> assignment -> loss -> gradients, against the reference's golden vectors."""
    loss, items, last, gs, gd, targets = _run_cuda(cuda_device, name, override=False)
    _check_against_golden(name, loss, items, last, gs, gd, targets, loss_rtol=1e-5, exact_boxes=False)
    assert loss.dtype == torch.float64 and items.shape == (3,)


@pytest.mark.parametrize("name", ["sparse", "crowded"])
@pytest.mark.parametrize("override", [True, False])
def test_warmup_atss_assigner(cuda_device, name, override):
    """The warm-up epochs (epoch_num < warmup_epoch): ATSS assignment (9 closest anchors per level, mean + std IoU threshold)
    with IoU soft labels, against the reference's golden vectors.  The foreground set and the assigned boxes are exact; the
    target scores are fp32 in the reference (rounded identically here); its class loss is an fp32 sum (float64 here): 1e-5."""
    loss, items, last, gs, gd, targets = _run_cuda(cuda_device, name, override=override, warmup=True)
    _check_against_golden(name + "_atss", loss, items, last, gs, gd, targets, loss_rtol=1e-5, exact_boxes=override,
                          score_rtol=1e-7 if override else 1e-4)


def test_detect_loss_matches_oracle_live(cuda_device):
    """Same comparison against the oracle run in this process (not only the stored vectors), incl. dense target scores."""
    name = "sparse"
    loss, items, last, gs, gd, targets = _run_cuda(cuda_device, name, override=True)
    scores, distri, _ = _losscases.make_case(name)
    ps, pd = scores.clone().requires_grad_(), distri.clone().requires_grad_()
    lo, io, asg = ol.compute_loss(ps, pd, targets, return_assignment=True)
    lo.backward()
    assert torch.equal(last["fg_mask"].cpu().bool(), asg["fg_mask"])
    f = asg["fg_mask"]
    assert torch.equal(last["target_gt_idx"].cpu().long()[f], asg["target_gt_idx"][f])
    np.testing.assert_allclose(last["target_score"].cpu().numpy(), asg["target_scores"].sum(-1).numpy(), rtol=1e-9, atol=0)
    np.testing.assert_allclose(loss.item(), lo.item(), rtol=1e-6)
    np.testing.assert_allclose(last["scalars"][4].item(), asg["target_scores_sum"].item(), rtol=1e-12)
    assert (gs.cpu() - ps.grad).abs().max() / ps.grad.abs().max() < 2e-4
    assert (gd.cpu() - pd.grad).abs().max() / pd.grad.abs().max() < 2e-4


def test_detect_loss_no_targets(cuda_device):
    """No box in the batch: the reference divides the class loss by target_scores_sum = 0 (inf) and returns 0 box terms."""
    loss, items, last, _, _, _ = _run_cuda(cuda_device, "empty", override=False)
    assert torch.isinf(loss) and items[0] == 0 and items[1] == 0 and torch.isinf(items[2])
    assert last["fg_mask"].sum() == 0


def test_detect_loss_capacity(cuda_device):
    """A larger per-image capacity changes nothing (padding rows never win); a smaller one is reported, not silent."""
    a = _run_cuda(cuda_device, "sparse", override=True, gt_cap=12, grads=False)
    b = _run_cuda(cuda_device, "sparse", override=True, gt_cap=40, grads=False)
    assert a[0].item() == b[0].item() and torch.equal(a[2]["fg_mask"], b[2]["fg_mask"])
    assert torch.equal(a[2]["target_gt_idx"][a[2]["fg_mask"].bool()], b[2]["target_gt_idx"][b[2]["fg_mask"].bool()])
    c = _run_cuda(cuda_device, "sparse", override=True, gt_cap=8, grads=False)
    assert int(c[2]["scalars"][6].item()) == 4  # image 2 has 12 boxes


def test_detect_loss_is_deterministic(cuda_device):
    a = _run_cuda(cuda_device, "crowded", override=False)
    b = _run_cuda(cuda_device, "crowded", override=False)
    assert a[0].item() == b[0].item() and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])


def test_no_cpu_path(cuda_device):
    from maf_yolo_b200.loss import ComputeLoss

    scores, distri, targets = _losscases.make_case("sparse")
    with pytest.raises(RuntimeError):  # no CPU path
        ComputeLoss(warmup_epoch=0)((None, scores, distri), targets, 0, 0)


@pytest.mark.parametrize("img,nc,warmup", [(320, 80, False), (640, 3, False), (320, 6, True), (64, 80, False)])
def test_other_image_sizes_and_class_counts(cuda_device, img, nc, warmup):
    """Image sizes other than 640 (anchor grid derived from the size), class counts that are not a multiple of 4 (the
    scalar varifocal path) and a tiny pyramid (64 px: 8 x 8 + 4 x 4 + 2 x 2 anchors, fewer than 9 cells on a level for the
    ATSS window) against the oracle run live."""
    from maf_yolo_b200.loss import ComputeLoss

    g = torch.Generator().manual_seed(100 + img + nc)
    a = (img // 8) ** 2 + (img // 16) ** 2 + (img // 32) ** 2
    bs = 2
    scores = torch.sigmoid(torch.randn(bs, a, nc, generator=g) * 1.5 - 2.0)
    distri = torch.randn(bs, a, 68, generator=g)
    rows = []
    for b in range(bs):
        for _ in range(4):
            cx, cy = (0.2 + 0.6 * torch.rand(2, generator=g)).tolist()
            w, h = (0.2 + 0.5 * torch.rand(2, generator=g)).tolist()
            rows.append([float(b), float(int(torch.randint(0, nc, (1,), generator=g))), cx, cy, w, h])
    targets = torch.tensor(rows, dtype=torch.float32)
    ps, pd = scores.clone().requires_grad_(), distri.clone().requires_grad_()
    lo, io, asg = ol.compute_loss(ps, pd, targets, img_size=img, num_classes=nc, return_assignment=True, warmup=warmup)
    lo.backward()
    crit = ComputeLoss(num_classes=nc, ori_img_size=img, warmup_epoch=3 if warmup else 0)
    cs, cd = scores.to(cuda_device).requires_grad_(), distri.to(cuda_device).requires_grad_()
    loss, items = crit((None, cs, cd), targets.to(cuda_device), 0, 0)
    loss.backward()
    torch.cuda.synchronize()
    assert torch.equal(crit.last["fg_mask"].cpu().bool(), asg["fg_mask"]), (int(crit.last["fg_mask"].sum()), int(asg["fg_mask"].sum()))
    f = asg["fg_mask"]
    assert int(f.sum()) > 0
    assert torch.equal(crit.last["target_gt_idx"].cpu().long()[f], asg["target_gt_idx"][f])
    np.testing.assert_allclose(loss.item(), lo.item(), rtol=2e-5)
    np.testing.assert_allclose(items.cpu().numpy(), io.double().numpy(), rtol=2e-5)
    assert (cs.grad.cpu() - ps.grad).abs().max() / ps.grad.abs().max() < 2e-4
    assert (cd.grad.cpu() - pd.grad).abs().max() / pd.grad.abs().max() < 2e-4
