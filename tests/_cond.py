"""Shared helpers for the conditioned end-to-end fixtures (tests/golden/cond_*.npz, made by make_golden_cond.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KEPT, SUP, UNK = 1, 2, 0


def pair_iou(b):
    """IoU matrix of xyxy boxes (float64; margins only)."""
    b = b.astype(np.float64)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    x1 = np.maximum(b[:, None, 0], b[None, :, 0])
    y1 = np.maximum(b[:, None, 1], b[None, :, 1])
    x2 = np.minimum(b[:, None, 2], b[None, :, 2])
    y2 = np.minimum(b[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    return inter / (area[:, None] + area[None, :] - inter)


def load_fixture(variant: str, kind: str):
    path = os.path.join(GOLD, f"cond_{variant}_{kind}.npz")
    if not os.path.exists(path):
        import pytest

        pytest.skip(f"{path} not generated (python tests/golden/make_golden_cond.py {variant})")
    return np.load(path)


def fixture_inputs(variant: str, fx):
    """(graph, state_dict, image) of a fixture, regenerated from its seeds (no reference needed)."""
    from maf_yolo_b200 import synth, topology
    from tests._synthetic import synthetic_scene

    g = topology.build_graph(variant)
    sd = synth.random_state_dict(g, seed=int(fx["seed"]), head_std=float(fx["head_std"]), cls_bias=float(fx["cls_bias"]),
                                 conv_gain=float(fx["conv_gain"]), reg_peak=float(fx["reg_peak"]),
                                 reg_sharp=float(fx["reg_sharp"]), reg_std=float(fx["reg_std"]))
    return g, sd, synthetic_scene(1, seed=int(fx["scene_seed"]))


def conditioned_inputs(variant: str, batch: int):
    """The conditioned weight family (the strict fixture's weights) with `batch` DIFFERENT scene images — for
    forward-parity tests at other batch sizes."""
    from tests._synthetic import synthetic_scene

    g, sd, _ = fixture_inputs(variant, load_fixture(variant, "strict"))
    return g, sd, synthetic_scene(batch, seed=11)


def match_detections(det: np.ndarray, fx, box_tol_px: float):
    """Matches each of OUR detections [n, 6] (xyxy, score, cls) to a reference candidate of the fixture (same class,
    every box coordinate within box_tol_px).  Returns the list of candidate indices (-1 = no match)."""
    cx, cc = fx["cand_xyxy"], fx["cand_cls"]
    out = []
    for row in det:
        ok = (cc == int(row[5])) & (np.abs(cx - row[None, :4]).max(1) <= box_tol_px)
        idx = np.nonzero(ok)[0]
        out.append(int(idx[np.argmin(np.abs(cx[idx] - row[None, :4]).max(1))]) if idx.size else -1)
    return out


def check_against_fixture(det: np.ndarray, fx, what: str, box_tol_px: float = 1.0):
    """The tolerance-aware set comparison the fixture certifies (see make_golden_cond.py):
      * every surely-kept reference detection appears in ours, score within margin_score / 2, box within box_tol_px;
      * none of ours is a surely-suppressed reference candidate, and every one of ours IS a reference candidate;
      * surely-kept detections whose reference scores differ by >= margin_score appear in the reference's order.
    With a STRICT fixture (no undecided / optional candidate) this is plain equality of the detection set with the
    reference's NMS output `det0` — also asserted directly (count, classes, order)."""
    st, opt, sc = fx["cand_status"], fx["cand_optional"], fx["cand_score"]
    m_s = float(fx["margin_score"])
    idx = match_detections(det, fx, box_tol_px)
    assert all(i >= 0 for i in idx), f"{what}: a detection of ours is not a reference candidate: {[det[k] for k, i in enumerate(idx) if i < 0][:3]}"
    assert len(set(idx)) == len(idx), f"{what}: two of our detections match the same reference candidate"
    assert not any(st[i] == SUP for i in idx), f"{what}: we kept a box the reference surely suppresses"
    must = set(np.nonzero((st == KEPT) & ~opt)[0].tolist())
    missing = must - set(idx)
    assert not missing, f"{what}: {len(missing)} surely-kept reference detections are missing from ours"
    for k, i in enumerate(idx):
        assert abs(float(det[k, 4]) - float(sc[i])) < m_s / 2, f"{what}: score of a matched detection is off by >= margin/2"
    ours_sure = [i for i in idx if i in must]
    for a in range(len(ours_sure)):
        for b in range(a + 1, len(ours_sure)):
            if abs(sc[ours_sure[a]] - sc[ours_sure[b]]) >= m_s:
                assert sc[ours_sure[a]] > sc[ours_sure[b]], f"{what}: output order differs from the reference's"
    strict = not ((st == UNK) | (opt & (st != SUP))).any()
    if strict:
        ref = fx["det0"]
        assert det.shape[0] == ref.shape[0], f"{what}: {det.shape[0]} detections, the reference has {ref.shape[0]}"
        assert np.array_equal(det[:, 5], ref[:, 5]), f"{what}: classes / order differ from the reference"
        assert np.abs(det[:, :4] - ref[:, :4]).max() <= box_tol_px and np.abs(det[:, 4] - ref[:, 4]).max() < m_s / 2
    return dict(strict=strict, n=det.shape[0], n_must=len(must),
                box_err=float(max(np.abs(fx["cand_xyxy"][i] - det[k, :4]).max() for k, i in enumerate(idx))),
                score_err=float(max(abs(float(det[k, 4]) - float(sc[i])) for k, i in enumerate(idx))))
