"""GPU parity of the block-level drop-ins (maf_yolo_b200.blocks): every block type of the reference that
is on the path, called with the reference block's forward signature (NCHW tensors), against the oracle's
restatement of that block (oracle/model.py, pinned to the reference)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    return (got.float().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)


def _setup(variant="n"):
    from maf_yolo_b200 import fold, synth, topology
    from oracle import model as om

    g = topology.build_graph(variant)
    sd = synth.random_state_dict(g, seed=0)
    return g, sd, fold.fold_state_dict(g, sd), om.parse_model(om.variant_rows(variant)), om


@pytest.mark.parametrize("layer,hw", [(0, 64), (1, 64), (2, 40), (3, 40), (4, 40), (9, 20), (10, 40), (12, 20), (31, 40), (33, 20)])
def test_block_matches_oracle(cuda_device, layer, hw):
    from maf_yolo_b200.blocks import B200Block

    g, sd, folded, spec, om = _setup()
    lay, d = g.layers[layer], spec[layer]
    p = f"backbone.{layer}"
    gen = torch.Generator().manual_seed(layer)
    if layer == 0:
        x = torch.rand(2, 3, hw, hw, generator=gen)
    else:
        x = torch.randn(2, sum(lay.c_in), hw, hw, generator=gen).half().float()
    with torch.no_grad():
        ref = {"repvgg": lambda: om._repvgg(sd, p, x), "rephdw": lambda: om._rephdw(sd, p, x, d),
               "mprep": lambda: om._mprep(sd, p, x), "sppf": lambda: om._sppf(sd, p, x),
               "convw": lambda: om._conv_mod(sd, p + ".block", x, 2), "head": lambda: om._head(sd, p, x, d)}[lay.kind]()
    blk = B200Block(g, folded, layer)
    assert blk.i == layer and blk.type.endswith(d["type"])
    # a RepHDW fed by a Concat receives the already concatenated tensor in the reference (yolo.py:193)
    got = blk(x.to(cuda_device))
    if lay.kind == "head":
        assert isinstance(got, tuple) and len(got) == 3
        for gt, rf, name in zip(got, ref, ("stem", "cls", "reg")):
            assert gt.shape == rf.shape and _rel(gt, rf) < 5e-3, (name, _rel(gt, rf))
        assert got[1].min() >= 0 and got[1].max() <= 1  # sigmoid applied, as Head_DepthUni.forward does
    else:
        assert got.shape == ref.shape and got.dtype == torch.float32
        assert _rel(got, ref) < 5e-3, _rel(got, ref)


def test_detect_dropin_matches_oracle(cuda_device):
    from maf_yolo_b200.blocks import B200Detect
    from oracle import model as om

    gen = torch.Generator().manual_seed(3)
    outs = []
    for c, hw in ((128, 16), (128, 8), (192, 4)):
        outs.append((torch.randn(2, c, hw, hw, generator=gen), torch.sigmoid(torch.randn(2, 80, hw, hw, generator=gen) - 3),
                     torch.randn(2, 68, hw, hw, generator=gen) * 2))
    ref = om.detect_eval(outs)
    det = B200Detect(80, [8, 16, 32])
    got = det([tuple(t.to(cuda_device) for t in lv) for lv in outs]).cpu()
    assert got.shape == ref.shape
    # the block boundary rounds the reg logits to fp16: norm-wise tolerance as in tests/test_model_gpu.py
    assert (got[..., :4] - ref[..., :4]).abs().max() <= 5e-4 * ref[..., :4].abs().max()
    assert (got[..., 5:] - ref[..., 5:]).abs().max() < 1e-3
    with pytest.raises(NotImplementedError):
        det(outs, val_loss=True)


def test_blocks_inside_the_reference_executor_loop(cuda_device):
    """The reference's Model.forward loop (yolo.py:189-209) with B200Block modules swapped in for every
    supported layer (what convert_blocks() does to a live reference model); Concat / Upsample / Out stay
    torch ops.  Compared with the oracle forward."""
    from maf_yolo_b200.blocks import SUPPORTED, B200Block, B200Detect

    g, sd, folded, spec, om = _setup()
    x = torch.rand(1, 3, 320, 320, generator=torch.Generator().manual_seed(0))
    ref = om.forward_train_form(spec, sd, x)
    blocks = {l.i: B200Block(g, folded, l.i) for l in g.layers if l.kind in SUPPORTED}
    y, cur = [], x.to(cuda_device)
    for d in spec:
        f = d["f"]
        if f != -1:
            cur = y[f] if isinstance(f, int) else [cur if j == -1 else y[j] for j in f]
        if d["i"] in blocks:
            cur = blocks[d["i"]](cur)
        elif d["type"] == "Concat":
            cur = torch.cat(cur, 1)
        elif d["type"] == "Upsample":
            cur = F.interpolate(cur, scale_factor=2.0, mode="nearest")
        elif d["type"] == "Out":
            cur = list(cur)
        y.append(cur)
    pred = B200Detect(80, g.strides)(cur).cpu()
    assert pred.shape == ref.shape
    box_err = (pred[..., :4] - ref[..., :4]).abs().max().item()
    assert box_err <= 1e-3 * ref[..., :4].abs().max().item(), box_err  # extra fp16 rounding at every block boundary
    assert (pred[..., 5:] - ref[..., 5:]).abs().max().item() <= 1e-3
