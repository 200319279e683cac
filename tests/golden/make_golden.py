"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The reference ships no tests, fixtures or golden vectors, so these are outputs of
the reference's own modules (yolov6.models.yolo.Model in eval mode, train form, and
yolov6.utils.nms.non_max_suppression) on seeded inputs:

  weights : maf_yolo_b200.synth.random_state_dict(graph, seed=0)   loaded with strict=True
  image   : torch.rand(B,3,640,640, generator=manual_seed(0))       fp32 in [0,1)
  nms     : eval settings conf 0.03 / iou 0.65 / multi_label / max_det 300 (yolov6/core/evaler.py:178)
            and demo settings conf 0.4 / iou 0.45 / single label / 1000 (tools/infer.py:24-26)

Files (all float32 unless noted):
  {v}_pred.npz     pred rows [::step] of image 0 (step 1 for n, 8 for s/m) + per-layer output statistics
  {v}_nms.npz      reference NMS output on the reference's own pred (eval settings)
  synth_nms.npz    reference NMS outputs on the synthetic score tensor of tests (seed 1)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from maf_yolo_b200 import synth, topology  # noqa: E402
from oracle import ref_loader  # noqa: E402

STEP = {"n": 1, "s": 8, "m": 8}


def synthetic_pred(b, a, nc, seed):
    """Same generator as tests/_synthetic.py::synthetic_pred."""
    from tests._synthetic import synthetic_pred as sp

    return sp(b, a, nc, seed)


def main():
    ns = ref_loader.load()
    torch.manual_seed(0)
    for v in "nsm":
        g = topology.build_graph(v)
        sd = synth.random_state_dict(g, seed=0)
        m = ref_loader.build_model(v)
        m.load_state_dict(sd, strict=True)
        x = torch.rand(2, 3, 640, 640, generator=torch.Generator().manual_seed(0))
        stats = {}
        hooks = []
        for i, mod in enumerate(m.backbone):
            def hook(_m, _inp, out, i=i):
                if isinstance(out, torch.Tensor):
                    o = out[0].float()
                    stats[i] = np.array([o.mean().item(), o.std().item(), o.abs().max().item()], dtype=np.float64)
            hooks.append(mod.register_forward_hook(hook))
        with torch.no_grad():
            pred = m(x)[0]
        for h in hooks:
            h.remove()
        layer_ids = np.array(sorted(stats), dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, f"{v}_pred.npz"), pred=pred[0, ::STEP[v]].numpy(),
                            step=np.int64(STEP[v]), layer_ids=layer_ids,
                            layer_stats=np.stack([stats[i] for i in layer_ids]))
        dets = ns.non_max_suppression(pred.clone(), 0.03, 0.65, multi_label=True)
        np.savez_compressed(os.path.join(HERE, f"{v}_nms.npz"), **{f"det{i}": d.numpy() for i, d in enumerate(dets)})
        print(v, "pred", tuple(pred.shape), "dets", [d.shape[0] for d in dets], "cand>0.03",
              int((pred[..., 5:] > 0.03).sum()))
    sp = synthetic_pred(3, 8400, 80, 1)
    out = {}
    for name, kw in (("eval", dict(conf_thres=0.03, iou_thres=0.65, multi_label=True)),
                     ("demo", dict(conf_thres=0.4, iou_thres=0.45, max_det=1000)),
                     ("low", dict(conf_thres=0.01, iou_thres=0.45, max_det=1000)),
                     ("agn", dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, agnostic=True, classes=[1, 5, 7]))):
        dets = ns.non_max_suppression(sp.clone(), **kw)
        for i, d in enumerate(dets):
            out[f"{name}_{i}"] = d.numpy()
        print("synth", name, [d.shape[0] for d in dets])
    np.savez_compressed(os.path.join(HERE, "synth_nms.npz"), **out)


if __name__ == "__main__":
    main()
