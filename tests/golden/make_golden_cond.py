"""Generates tests/golden/cond_{v}.npz: CONDITIONED end-to-end fixtures from the UNMODIFIED reference.

Why.  With PyTorch-default init (make_golden.py's weights) a signal decays ~0.4x per layer, so after 30+ layers the heads
see almost only the last layers' biases: predictions barely depend on the image, ~900 candidates per image are nearly
tied and overlap near the IoU threshold, and the reference's detection SET is not a stable property of its input (its
own fp16 mode reshuffles it).  Here:
  weights : synth.random_state_dict(conv_gain=2.15; 2.05 for M): He-like scale, activations keep their magnitude through the whole
            network (every layer's error reaches the output, as with trained weights); DFL head shaped to ~16-cell boxes
            (reg_peak / reg_sharp / reg_std); class bias set so that the image's largest class logit is PEAK_LOGIT
  image   : tests/_synthetic.synthetic_scene — structure at every anchor scale, no flat regions
  output  : the reference's own Model.forward (train form, eval mode, fp32 CPU) -> non_max_suppression
            (yolov6/utils/nms.py:31, multi_label=True)

Greedy NMS is discontinuous, so "same detections within fp tolerance" needs a definition.  The generator CERTIFIES, per
candidate, whether the reference's decision is stable under ANY perturbation of every score by < M_SCORE / 2 and of
every pair IoU by < M_IOU (three-valued greedy NMS, `certify`):
  status 1  surely kept        must appear in our output (class, box, score within tolerance), in the reference's order
  status 2  surely suppressed  must NOT appear in our output
  status 0  undecided          the reference's own answer flips inside the tolerance: either outcome is accepted
  optional  score within M_SCORE / 2 of conf: may or may not be a candidate at all
and searches (seed, conf) for fixtures with many status-1 and status-2 candidates and few undecided ones; a second,
STRICT fixture per variant has no undecided / optional candidate at all, so there the detection set must be EQUAL.

Run in the build container:  python tests/golden/make_golden_cond.py [n s m]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from maf_yolo_b200 import synth, topology  # noqa: E402
from oracle import ref_loader  # noqa: E402
from tests._cond import pair_iou  # noqa: E402
from tests._synthetic import synthetic_scene  # noqa: E402

# margins: >= 2x the score error measured on B200 for the variant's fixtures (profiles/r02_b_parity_k7on.txt: N 2.2e-3,
# S 8.0e-3; deeper variants amplify more), IoU margin far above the measured 1.5e-4
M_SCORE_V = {"n": 1.0e-2, "s": 2.0e-2, "m": 2.0e-2}
M_SCORE, M_IOU = 1.0e-2, 0.05
IOU_THRES = 0.65
MAX_CAND = 600
HEAD_STD, CLS_BIAS0 = 0.5, -5.0
CONV_GAIN = {"n": 2.15, "s": 2.15, "m": 2.05}  # M is deeper (2-4 bottlenecks per stage): 2.15 overflows for most seeds
REG = dict(reg_peak=8.0, reg_sharp=0.3, reg_std=0.1)
PEAK_LOGIT = 1.5
SCENE_SEED = 7
KEPT, SUP, UNK = 1, 2, 0
RICH_MIN = {"n": 15, "s": 15, "m": 8}      # surely-kept / surely-suppressed candidates a "rich" fixture needs
RICH_LOOSE = {"n": 0.3, "s": 0.3, "m": 0.45}
STRICT_MIN = {"n": 3, "s": 3, "m": 2}


def certify(pred_img, conf, m_score=M_SCORE, m_iou=M_IOU):
    """Three-valued greedy NMS of one image's [A, 85] prediction at `conf` (multi-label, class-offset boxes):
      surely suppressed : a surely-kept, surely-present candidate of the same class is clearly ahead (score gap >=
                          m_score) and clearly overlapping (IoU >= thr + m_iou)
      surely kept       : surely present, and every same-class candidate that COULD precede it (score > own - m_score)
                          is clearly not overlapping (IoU <= thr - m_iou) or surely suppressed (a suppressed box
                          suppresses nobody)
    iterated to a fixed point.  Returns None if there are too many candidates, else a dict of per-candidate arrays
    (everything with score > conf - m_score / 2): anchor, cls, score, xyxy, status, optional."""
    sc = pred_img[:, 5:]
    a_idx, c_idx = np.nonzero(sc > conf - m_score / 2)
    n = a_idx.shape[0]
    if n == 0 or n > MAX_CAND:
        return None
    s = sc[a_idx, c_idx].astype(np.float64)
    optional = s < conf + m_score / 2
    xywh = pred_img[a_idx, :4]
    xyxy = np.concatenate([xywh[:, :2] - xywh[:, 2:] / 2, xywh[:, :2] + xywh[:, 2:] / 2], 1)
    iou = pair_iou(xyxy)
    same = (c_idx[:, None] == c_idx[None, :]) & ~np.eye(n, dtype=bool)
    clearly_over = same & (iou >= IOU_THRES + m_iou)
    maybe_over = same & (iou > IOU_THRES - m_iou)
    clearly_ahead = (s[None, :] - s[:, None]) >= m_score      # [i, j]: j clearly ahead of i
    maybe_ahead = (s[None, :] - s[:, None]) > -m_score        # [i, j]: j could precede i
    st = np.zeros(n, dtype=np.int64)
    order = np.argsort(-s)
    for _ in range(n + 2):
        changed = False
        for i in order:
            if st[i] != UNK:
                continue
            if (clearly_over[i] & clearly_ahead[i] & (st == KEPT) & ~optional).any():
                st[i] = SUP; changed = True
            elif not optional[i] and not (maybe_over[i] & maybe_ahead[i] & (st != SUP)).any():
                st[i] = KEPT; changed = True
        if not changed:
            break
    return dict(anchor=a_idx.astype(np.int64), cls=c_idx.astype(np.int64), score=s.astype(np.float32),
                xyxy=xyxy.astype(np.float32), status=st, optional=optional)


def cond_state_dict(g, m, x, seed, gain):
    """Seeded weights of the conditioned family; the class bias is then set (one extra reference forward) so that the
    largest class logit of the image is PEAK_LOGIT: the top scores sit where the sigmoid is steep.  Returns (sd, bias)
    or None if the seed's activations overflow.  synth.random_state_dict(..., cls_bias=bias) reproduces sd exactly."""
    sd = synth.random_state_dict(g, seed=seed, head_std=HEAD_STD, cls_bias=CLS_BIAS0, conv_gain=gain, **REG)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        p = m(x)[0]
    mx = torch.logit(p[..., 5:].double().clamp(1e-12, 1 - 1e-12)).max().item()
    if not np.isfinite(mx) or mx > 12:
        return None
    bias = round(CLS_BIAS0 - (mx - PEAK_LOGIT), 3)
    return synth.random_state_dict(g, seed=seed, head_std=HEAD_STD, cls_bias=bias, conv_gain=gain, **REG), bias


def search(variant, ns, seeds=range(380, 440)):
    """Returns (rich, strict): the best fixture by number of surely-kept detections (>= 15 surely kept, >= 15 surely
    suppressed, undecided + optional <= 30 % of the candidates), and the best fixture without any undecided / optional
    candidate (detection set must be equal)."""
    g = topology.build_graph(variant)
    m = ref_loader.build_model(variant)
    x = synthetic_scene(1, seed=SCENE_SEED)
    rich = strict = None
    for seed in seeds:
        r = cond_state_dict(g, m, x, seed, CONV_GAIN[variant])
        if r is None:
            continue
        sd, bias = r
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            pred = m(x)[0]
        p = pred.numpy()[0]
        top = np.sort(p[:, 5:].reshape(-1))[::-1][:MAX_CAND + 1]
        for k in range(4, MAX_CAND, 2):
            conf = float((top[k - 1] + top[k]) / 2)
            c = certify(p, conf, M_SCORE_V[variant])
            if c is None:
                continue
            nk, nsup = int((c["status"] == KEPT).sum()), int((c["status"] == SUP).sum())
            loose = int(((c["status"] == UNK) | (c["optional"] & (c["status"] != SUP))).sum())
            cand = dict(seed=seed, bias=bias, conf=conf, pred=pred, cert=c, nk=nk, nsup=nsup, loose=loose)
            if loose == 0 and nk >= STRICT_MIN[variant] and (strict is None or nk + nsup > strict["nk"] + strict["nsup"]):
                strict = cand
            if (nk >= RICH_MIN[variant] and nsup >= RICH_MIN[variant] and loose <= RICH_LOOSE[variant] * len(c["status"])
                    and (rich is None or nk > rich["nk"])):
                rich = cand
        print(f"  {variant}: seed {seed} (bias {bias}) -> rich {None if rich is None else (rich['seed'], rich['nk'], rich['nsup'], rich['loose'])}"
              f" strict {None if strict is None else (strict['seed'], strict['nk'], strict['nsup'])}", flush=True)
    return rich, strict


def save(variant, kind, fx, ns):
    dets = ns.non_max_suppression(fx["pred"].clone(), fx["conf"], IOU_THRES, multi_label=True)
    c = fx["cert"]
    out = {"seed": np.int64(fx["seed"]), "conf": np.float64(fx["conf"]), "iou": np.float64(IOU_THRES),
           "head_std": np.float64(HEAD_STD), "cls_bias": np.float64(fx["bias"]), "conv_gain": np.float64(CONV_GAIN[variant]),
           "scene_seed": np.int64(SCENE_SEED), **{k: np.float64(v) for k, v in REG.items()},
           "margin_score": np.float64(M_SCORE_V[variant]), "margin_iou": np.float64(M_IOU),
           "det0": dets[0].numpy(), "pred_sample": fx["pred"][0, ::16].numpy(),
           **{"cand_" + k: v for k, v in c.items()}}
    np.savez_compressed(os.path.join(HERE, f"cond_{variant}_{kind}.npz"), **out)
    print(variant, kind, "seed", fx["seed"], "conf %.4f" % fx["conf"], "candidates", len(c["status"]), "surely kept", fx["nk"],
          "surely suppressed", fx["nsup"], "undecided/optional", fx["loose"], "reference kept", dets[0].shape[0])


def main():
    ns = ref_loader.load()
    for v in sys.argv[1:] or "nsm":
        rich, strict = search(v, ns)
        if rich is None or strict is None:
            raise SystemExit(f"{v}: no fixture found (rich {rich is not None}, strict {strict is not None})")
        save(v, "rich", rich, ns)
        save(v, "strict", strict, ns)


if __name__ == "__main__":
    main()
