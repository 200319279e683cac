"""CPU tests of the oracle (oracle/) against the committed golden vectors, which were produced by the
unmodified reference (tests/golden/make_golden.py), and — when /root/reference is present — against
the reference itself.  The oracle is the checker for every GPU parity test, so it is pinned first."""
import os

import numpy as np
import pytest
import torch

from maf_yolo_b200 import synth, topology
from oracle import model as om
from oracle import nms as onms
from oracle import ref_loader
from tests._synthetic import synthetic_image, synthetic_pred

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _oracle_pred(variant, batch=1):
    g = topology.build_graph(variant)
    sd = synth.random_state_dict(g, seed=0)
    spec = om.parse_model(om.variant_rows(variant))
    x = synthetic_image(2, seed=0)[:batch]
    return om.forward_train_form(spec, sd, x), spec, sd, x


@pytest.mark.parametrize("variant", ["n", "s", "m"])
def test_oracle_forward_matches_golden(variant):
    gold = np.load(os.path.join(GOLD, f"{variant}_pred.npz"))
    pred, _, _, _ = _oracle_pred(variant)
    got = pred[0, ::int(gold["step"])].numpy()
    ref = gold["pred"]
    assert got.shape == ref.shape
    # same fp32 ops in the same order: identical up to oneDNN thread-count effects
    np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=0, atol=2e-3)
    np.testing.assert_allclose(got[:, 4:], ref[:, 4:], rtol=0, atol=1e-6)


def test_oracle_layer_stats_match_golden():
    gold = np.load(os.path.join(GOLD, "n_pred.npz"))
    ids = [int(i) for i in gold["layer_ids"]]
    g = topology.build_graph("n")
    sd = synth.random_state_dict(g, seed=0)
    spec = om.parse_model(om.variant_rows("n"))
    _, kept = om.forward_train_form(spec, sd, synthetic_image(2, seed=0), keep=ids)
    for row, i in zip(gold["layer_stats"], ids):
        o = kept[i][0].float()
        got = np.array([o.mean().item(), o.std().item(), o.abs().max().item()])
        np.testing.assert_allclose(got, row, rtol=1e-4, atol=1e-5, err_msg=f"layer {i}")


@pytest.mark.parametrize("variant", ["n", "s"])
def test_oracle_deploy_fold_equals_train_form(variant):
    """The deploy conversion is exact in real arithmetic: fp32 noise only (1.2e-4 px)."""
    pred, spec, sd, x = _oracle_pred(variant)
    dep = om.forward_deploy(spec, om.fold_deploy(spec, sd), x)
    assert (dep[..., :4] - pred[..., :4]).abs().max().item() < 2e-3
    assert (dep[..., 5:] - pred[..., 5:]).abs().max().item() < 1e-6


def test_oracle_nms_matches_golden():
    gold = np.load(os.path.join(GOLD, "synth_nms.npz"))
    sp = synthetic_pred(3, 8400, 80, 1).numpy()
    cases = dict(eval=dict(conf_thres=0.03, iou_thres=0.65, multi_label=True),
                 demo=dict(conf_thres=0.4, iou_thres=0.45, max_det=1000),
                 low=dict(conf_thres=0.01, iou_thres=0.45, max_det=1000),
                 agn=dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, agnostic=True, classes=[1, 5, 7]))
    for name, kw in cases.items():
        out = onms.non_max_suppression(sp, **kw)
        for i, d in enumerate(out):
            assert np.array_equal(d, gold[f"{name}_{i}"]), f"{name} image {i}: oracle NMS differs from the reference"


def test_oracle_nms_on_golden_pred_is_bitexact():
    """NMS bit-exactness is only meaningful on IDENTICAL inputs: feed the reference's
    own pred (golden, image 0 of MAF-YOLO-N) to the oracle NMS and compare with the reference's NMS output."""
    pred = np.load(os.path.join(GOLD, "n_pred.npz"))["pred"][None]
    ref = np.load(os.path.join(GOLD, "n_nms.npz"))["det0"]
    out = onms.non_max_suppression(pred, 0.03, 0.65, multi_label=True)[0]
    assert out.shape == ref.shape and ref.shape[0] > 50
    assert np.array_equal(out, ref)


def test_oracle_nms_edge_cases():
    """Empty input, single box, all-identical boxes, exact-threshold IoU (strict '>')."""
    z = np.zeros((1, 16, 85), dtype=np.float32)
    assert onms.non_max_suppression(z)[0].shape == (0, 6)
    one = z.copy()
    one[0, 3] = [10, 10, 4, 4, 1.0] + [0.0] * 80
    one[0, 3, 5 + 7] = 0.9
    d = onms.non_max_suppression(one)[0]
    assert d.shape == (1, 6) and d[0, 5] == 7 and np.allclose(d[0, :4], [8, 8, 12, 12])
    same = z.copy()
    same[0, :, :5] = [50, 50, 20, 20, 1.0]
    same[0, :, 5] = 0.5
    assert onms.non_max_suppression(same)[0].shape == (1, 6)
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 5]], dtype=np.float32)  # IoU exactly 0.5
    sc = np.array([0.9, 0.8], dtype=np.float32)
    assert list(onms.nms_indices(boxes, sc, 0.5)) == [0, 1]
    assert list(onms.nms_indices(boxes, sc, 0.4999)) == [0]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_oracle_bitexact_vs_reference_modules():
    """Strongest pin: oracle == the reference's own Model.forward / non_max_suppression, bit for bit."""
    ns = ref_loader.load()
    g = topology.build_graph("n")
    sd = synth.random_state_dict(g, seed=3)
    m = ref_loader.build_model("n")
    m.load_state_dict(sd, strict=True)
    x = synthetic_image(1, seed=5)
    with torch.no_grad():
        ref = m(x)[0]
    spec = om.parse_model(om.variant_rows("n"))
    assert torch.equal(om.forward_train_form(spec, sd, x), ref)
    sp = synthetic_pred(2, 8400, 80, 11)
    for kw in (dict(conf_thres=0.03, iou_thres=0.65, multi_label=True), dict(conf_thres=0.02, iou_thres=0.45)):
        for r, o in zip(ns.non_max_suppression(sp.clone(), **kw), onms.non_max_suppression(sp.numpy(), **kw)):
            assert np.array_equal(r.numpy(), o)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (GPU box)")
def test_oracle_train_form_outputs_vs_reference():
    """`Model.forward(x, val_loss=True)` in eval mode (yolo.py:333-354): (feats, cls_score_list, reg_distri_list) — the
    oracle's undecoded forward, flattened the same way, bit for bit."""
    g = topology.build_graph("n")
    sd = synth.random_state_dict(g, seed=3)
    m = ref_loader.build_model("n")
    m.load_state_dict(sd, strict=True)
    x = synthetic_image(1, seed=5)
    with torch.no_grad():
        (feats, scores, distri), _ = m(x, val_loss=True)
    spec = om.parse_model(om.variant_rows("n"))
    outs = om.forward_train_form(spec, sd, x, decode=False)
    assert torch.equal(scores, torch.cat([o[1].flatten(2).permute(0, 2, 1) for o in outs], 1))
    assert torch.equal(distri, torch.cat([o[2].flatten(2).permute(0, 2, 1) for o in outs], 1))
    assert all(torch.equal(a, o[0]) for a, o in zip(feats, outs))


@pytest.mark.parametrize("variant", ["n", "s", "m"])
@pytest.mark.parametrize("kind", ["strict", "rich"])
def test_oracle_reproduces_conditioned_fixtures(variant, kind):
    """The conditioned end-to-end fixtures (tests/golden/make_golden_cond.py: reference Model.forward ->
    reference non_max_suppression on He-scaled weights and a structured image) are reproduced bit for bit by the
    oracle forward + oracle NMS, from the seeds alone."""
    from tests import _cond

    if variant != "n" and os.environ.get("MAFB200_SLOW_CPU_TESTS", "1") == "0":
        pytest.skip("slow")
    fx = _cond.load_fixture(variant, kind)
    g, sd, x = _cond.fixture_inputs(variant, fx)
    spec = om.parse_model(om.variant_rows(variant))
    pred = om.forward_train_form(spec, sd, x)
    dets = onms.non_max_suppression(pred.numpy(), float(fx["conf"]), float(fx["iou"]), multi_label=True)
    if ref_loader.available():  # the machine that generated the fixtures: bit for bit
        assert np.array_equal(pred[0, ::16].numpy(), fx["pred_sample"])
        assert np.array_equal(dets[0], fx["det0"])
    else:  # another host CPU: torch's conv kernels differ in the last bits
        assert np.allclose(pred[0, ::16].numpy(), fx["pred_sample"], rtol=2e-4, atol=2e-5)
        _cond.check_against_fixture(dets[0], fx, f"{variant}/{kind} oracle on this host")
    # the certification stored with the fixture is consistent with the reference's own output
    st, opt = fx["cand_status"], fx["cand_optional"]
    r = _cond.check_against_fixture(fx["det0"], fx, f"{variant}/{kind} reference vs its own certification")
    assert r["strict"] == (kind == "strict") and r["n_must"] == int(((st == 1) & ~opt).sum())


def test_torchvision_form_of_the_nms_oracle_equals_the_numpy_form():
    """bench.py times `non_max_suppression_tv` (torchvision.ops.nms, the library call of nms.py:96) as the reference's
    NMS; it must return exactly what the numpy restatement (the parity checker) returns."""
    pred = synthetic_pred(2, 8400, 80, seed=1)
    for kw in (dict(conf_thres=0.03, iou_thres=0.65, multi_label=True), dict(conf_thres=0.4, iou_thres=0.45, max_det=1000),
               dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, agnostic=True, classes=[1, 5, 7])):
        a = onms.non_max_suppression(pred.numpy(), **kw)
        b = onms.non_max_suppression_tv(pred.clone(), **kw)
        assert all(np.array_equal(x, y.numpy()) for x, y in zip(a, b)), kw
