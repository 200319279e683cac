"""CPU tests of the host side: the C-ABI library loads and exports what include/mafb200.h declares,
argument validation works without a GPU, and the Python host logic (topology, fold, packing, plan,
arena) is right.  No compute kernels are launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from maf_yolo_b200 import _lib, engine, fold, ops, synth, topology
from oracle import model as om

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mafb200.h")).read()
    declared = set(re.findall(r"MAFB200_API\s+[\w\s\*]+?\b(mafb200_\w+)\s*\(", header))
    assert len(declared) >= 19
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    lib = _lib.lib()
    for name in declared:
        assert hasattr(lib, name), f"libmafb200.so does not export {name}"
    assert lib.mafb200_version() == 200


def test_tiling_and_packed_k():
    # every cout of the N/S/M graphs splits into <=256-wide, 16-aligned tiles that cover it
    for cout in [24, 48, 68, 72, 80, 96, 128, 144, 192, 256, 288, 384, 512, 576, 768, 1152, 1536]:
        nt, tn = _lib.gemm_tiling(cout)
        assert tn % 16 == 0 and 16 <= tn <= 128 and nt * tn >= cout and (nt - 1) * tn < cout
    assert _lib.gemm_tiling(288) == (3, 96) and _lib.gemm_tiling(24) == (1, 32) and _lib.gemm_tiling(192) == (2, 96)
    assert _lib.packed_k_1x1([96, 384]) == 128 + 384 and _lib.packed_k_1x1([24, 24, 24]) == 192
    assert _lib.packed_k_3x3(24) == 576 and _lib.packed_k_3x3(192) == 1728


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any device work; on this GPU-less machine a well-formed
    call fails with MAF_E_ARCH — never a crash, never a silent CPU fallback."""
    lib = _lib.lib()
    t = _lib.MafTensor(0x1000, 1, 8, 8, 16, 16, _lib.MAF_F16)
    bad = _lib.MafTensor(0x1000, 1, 8, 8, 16, 8, _lib.MAF_F16)  # c_stride < c
    assert lib.mafb200_conv1x1(None, 1, None, None, 0, C.byref(t), None, None) == -1
    arr = (_lib.MafTensor * 1)(bad)
    assert lib.mafb200_conv1x1(arr, 1, 0x1000, 0x1000, 1, C.byref(t), None, None) == -1
    assert b"src[0]" in lib.mafb200_last_error()
    arr = (_lib.MafTensor * 1)(t)
    assert lib.mafb200_conv1x1(arr, 1, 0x1000, 0x1000, 9, C.byref(t), None, None) == -1  # bad act
    assert lib.mafb200_dwconv(C.byref(t), 0x1000, 0x1000, 4, 0, C.byref(t), None) in (-1, -3)
    assert lib.mafb200_nms(0x1000, 1, 10, 80, 1.5, 0.5, 0, 0, None, 300, 30000, 0x1000, 0x1000, 0x1000, 1 << 30,
                           None) == -1
    assert b"conf_thres" in lib.mafb200_last_error()
    if not torch.cuda.is_available():
        rc = lib.mafb200_conv1x1(arr, 1, 0x1000, 0x1000, 1, C.byref(t), None, None)
        assert rc == -3 and b"CUDA" in lib.mafb200_last_error()
        with pytest.raises(_lib.MafError):
            _lib.check(rc)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            engine.Engine(topology.build_graph("n"), {}, 1)
    assert lib.mafb200_nms_workspace_bytes(32, 8400, 80) == 256 + 32 * (1 << 20) * 8 + 256


@pytest.mark.parametrize("variant,params_m,gflops", [("n", 3.761, 10.508), ("s", 8.554, 25.448), ("m", 23.697, 76.650)])
def test_topology_reproduces_readme_counts(variant, params_m, gflops):
    """Deploy-form parameter / FLOP counts of the resolved graph equal the reference README table
    (README.md:24-26) — the cheapest whole-topology check there is."""
    g = topology.build_graph(variant)
    folded = fold.fold_state_dict(g, synth.random_state_dict(g, 0))
    assert abs(fold.deploy_param_count(folded) / 1e6 - params_m) < 5e-4
    plan = engine.Plan(g, 640, 640)
    assert abs(plan.flops_per_image() / 1e9 - gflops) < 2e-3  # the 3 DFL proj convs (17->1) are the difference
    assert plan.anchors == 8400
    # product topology == oracle's restatement of parse_model
    spec = om.parse_model(om.variant_rows(variant))
    assert [l.c_out for l in g.layers] == [d["c2"] for d in spec]
    assert g.save == om.savelist(spec)


def test_fold_matches_oracle_and_accepts_deploy_form():
    g = topology.build_graph("n")
    sd = synth.random_state_dict(g, 1)
    spec = om.parse_model(om.variant_rows("n"))
    dd = om.fold_deploy(spec, sd, dtype=torch.float64)
    fa = fold.fold_state_dict(g, sd)
    key_map = {}
    deploy_sd = {}
    for k, (w, b) in dd.items():
        parts = k.split(".")
        leaf = parts[-1]
        short = ".".join(parts[1:])
        if leaf in ("cls_pred", "reg_pred"):
            deploy_sd[k + ".weight"], deploy_sd[k + ".bias"] = w, b
        elif (leaf == "conv2" and len(parts) == 5) or leaf in ("cls_conv", "reg_conv"):
            deploy_sd[k + ".dwconv.lk_origin.weight"], deploy_sd[k + ".dwconv.lk_origin.bias"] = w, b
            short = short.replace(".conv2", ".dw").replace("cls_conv", "cls_dw").replace("reg_conv", "reg_dw")
        elif w.shape[-1] == 3 and leaf != "block":
            deploy_sd[k + ".rbr_reparam.weight"], deploy_sd[k + ".rbr_reparam.bias"] = w, b
        else:
            deploy_sd[k + ".conv.weight"], deploy_sd[k + ".conv.bias"] = w, b
        short = short.replace("cls_conv_s", "cls_s").replace("reg_conv_s", "reg_s")
        key_map[short] = k
    assert set(key_map) == set(fa), set(key_map) ^ set(fa)
    for short, k in key_map.items():
        assert torch.allclose(fa[short][0], dd[k][0], rtol=1e-10, atol=1e-12), short
        assert torch.allclose(fa[short][1], dd[k][1], rtol=1e-10, atol=1e-12), short
    fb = fold.fold_state_dict(g, deploy_sd)  # deploy-form input (what a fused reference model holds)
    for short in fa:
        assert torch.allclose(fa[short][0], fb[short][0], rtol=1e-10, atol=1e-12), short
        assert torch.allclose(fa[short][1], fb[short][1], rtol=1e-10, atol=1e-12), short
    with pytest.raises(KeyError, match="neither train-form nor deploy-form"):
        fold.fold_state_dict(g, {"backbone.0.foo": torch.zeros(1)})


def test_weight_packing_layout():
    w = torch.arange(24 * 100, dtype=torch.float32).reshape(24, 100) / 1000
    b = torch.arange(24, dtype=torch.float32)
    wp, bp = ops.pack_conv1x1(w, b, [36, 64], device="cpu")
    assert wp.shape == (32, 64 + 64) and wp.dtype == torch.float16 and bp.shape == (32,)
    assert torch.equal(wp[:24, :36], w[:, :36].half()) and torch.equal(wp[:24, 64:128], w[:, 36:].half())
    assert (wp[:, 36:64] == 0).all() and (wp[24:] == 0).all() and (bp[24:] == 0).all()
    w3 = torch.randn(48, 24, 3, 3)
    wp3, _ = ops.pack_conv3x3(w3, torch.zeros(48), device="cpu")
    assert wp3.shape == (48, 9 * 64)
    assert torch.equal(wp3[:, 4 * 64: 4 * 64 + 24], w3[:, :, 1, 1].half())  # centre tap (ky=1,kx=1) is block 4
    wd, _ = ops.pack_dw(torch.randn(16, 1, 5, 5), torch.zeros(16), device="cpu")
    assert wd.shape == (5, 5, 16)
    ws, _ = ops.pack_stem(torch.randn(24, 3, 3, 3), torch.zeros(24), device="cpu")
    assert ws.shape == (24, 3, 3, 3)


@pytest.mark.parametrize("variant,bneck", [("n", "0"), ("m", "0"), ("n", "1"), ("s", "1")])
def test_plan_liveness_and_arena(variant, bneck, monkeypatch):
    """No two buffers that are live at the same launch overlap in the arena; slices stay in bounds.  bneck = "1": the
    plan with the fused DepthBottleneckUni kernel (K4, opt-in)."""
    monkeypatch.setenv("MAFB200_BNECK", bneck)
    g = topology.build_graph(variant)
    plan = engine.Plan(g, 640, 640)
    batch = 4
    total = plan.assign_offsets(batch)
    naive = sum(b.nbytes(batch) for b in plan.bufs)
    assert total < 0.45 * naive
    for b in plan.bufs:
        assert b.first is not None and b.first <= b.last and b.offset % 1024 == 0
        assert b.offset + b.nbytes(batch) <= total
    for i, a in enumerate(plan.bufs):
        for b in plan.bufs[i + 1:]:
            if a.first <= b.last and b.first <= a.last:  # live ranges intersect
                assert a.offset + a.nbytes(batch) <= b.offset or b.offset + b.nbytes(batch) <= a.offset, (a.name, b.name)
    for op in plan.ops:
        for v in op.reads + op.writes:
            assert v.c_off + v.c <= v.buf.ld and v.c_off % 8 == 0
    # every weighted op has folded weights of the right shape
    folded = fold.fold_state_dict(g, synth.random_state_dict(g, 0))
    for op in plan.ops:
        if op.weight:
            w, bias = folded[op.weight]
            if op.kind == "dwpw":  # fused depth-wise + 1x1: depth-wise weight over the INPUT channels, 1x1 to the output
                w2, b2 = folded[op.weight2]
                assert w.shape[0] == op.reads[0].c == bias.shape[0] and w2.shape[0] == op.writes[0].c == b2.shape[0], op.name
                assert w2.reshape(w2.shape[0], -1).shape[1] == op.reads[0].c, op.name
                continue
            if op.kind == "bneck":  # K4: expand 1x1 (c_ -> mid), depth-wise (mid), project 1x1 (mid -> c_)
                wd, bd = folded[op.weight3]
                w2, b2 = folded[op.weight2]
                mid = w.shape[0]
                assert w.reshape(mid, -1).shape[1] == op.reads[0].c and wd.shape[0] == mid == bd.shape[0], op.name
                assert w2.shape[0] == op.writes[0].c == b2.shape[0] and w2.reshape(w2.shape[0], -1).shape[1] == mid, op.name
                continue
            if op.kind == "head_pred":  # K7: no arena output; cls_pred -> nc scores, reg_pred -> 4 * 17 bins
                assert not op.writes and w.shape[0] == bias.shape[0] == (g.nc if op.act == "cls" else 68), op.name
                assert w.reshape(w.shape[0], -1).shape[1] == op.reads[0].c, op.name
                continue
            if op.wslice is not None:  # a split depth-wise conv uses a channel range of the folded weight
                w, bias = w[op.wslice[0]:op.wslice[1]], bias[op.wslice[0]:op.wslice[1]]
            assert w.shape[0] == op.writes[0].c == bias.shape[0], op.name
            if op.kind == "conv1x1":
                assert w.shape[1] == sum(v.c for v in op.reads), op.name
    # K7: six head-prediction launches carry the decode; anchor offsets tile [0, 8400); the counter reset precedes them
    hp = [op for op in plan.ops if op.kind == "head_pred"]
    assert plan.k7 and len(hp) == 6 and not any(op.kind == "decode" for op in plan.ops)
    assert sorted({(op.level, op.anchor_off) for op in hp}) == [(0, 0), (1, 6400), (2, 8000)] and plan.anchors == 8400
    kinds = [op.kind for op in plan.ops]
    assert kinds.count("detect_reset") == 1 and kinds.index("detect_reset") < kinds.index("head_pred")
    old = engine.Plan(g, 640, 640, k7=False)
    assert not old.k7 and sum(op.kind == "decode" for op in old.ops) == 1 and len(old.ops) == len(plan.ops)
    if bneck == "1":  # K4: every k <= 5, c_ <= 64 bottleneck is one kernel and its 3c_-wide buffer is gone
        names = [op.name for op in plan.ops if op.kind == "bneck"]
        assert names and not any(b.name.endswith(".expand") and b.name.split(".")[0] in {n.split(".")[0] for n in names}
                                 for b in plan.bufs)
        if variant == "n":
            assert names == ["L2.m0.bottleneck(k3)", "L4.m0.bottleneck(k5)", "L20.m0.bottleneck(k5)", "L22.m0.bottleneck(k5)"]
            assert plan.bytes_per_image() < 135e6
    else:
        assert not any(op.kind == "bneck" for op in plan.ops)
    # MAFPN fusion concats became multi-source GEMMs, upsample was fused away
    assert not any(op.kind == "upsample2x" for op in plan.ops)
    assert max(len(op.reads) for op in plan.ops if op.kind == "conv1x1") == 4


@pytest.mark.parametrize("variant", ["n", "s", "m"])
def test_pad_fill_only_touches_unowned_padding(variant, monkeypatch):
    monkeypatch.setenv("MAFB200_BNECK", "1")
    """Ops that end inside a 32-byte sector at the end of a padded buffer get zero filters for the padding channels
    (whole-sector stores).  Those channels must belong to no other view, the op must be the buffer's last-channel
    writer, and sigmoid outputs (act(0) != 0) are never extended."""
    monkeypatch.delenv("MAFB200_PAD_FILL", raising=False)
    plan = engine.Plan(topology.build_graph(variant), 640, 640)
    extended = {}
    for op in plan.ops:
        if not op.writes or op.kind not in ("stem", "conv1x1", "dwpw", "poolpw", "bneck"):
            continue
        act = op.act2 if op.kind in ("dwpw", "bneck") else op.act
        extra = engine.pad_fill_channels(op, act)
        v = op.writes[0]
        if extra:
            assert act != "sigmoid" and len(op.writes) == 1
            assert v.c_off + v.c == v.buf.c and v.c_off + v.c + extra == v.buf.ld
            assert ((v.c_off + v.c) * 2) % 32 != 0 and (v.buf.ld * 2) % 32 == 0
            extended[op.name] = (id(v.buf), v.buf.c, v.buf.ld)
        else:
            assert ((v.c_off + v.c) * 2) % 32 == 0 or v.c_off + v.c != v.buf.c or len(op.writes) != 1 or act == "sigmoid"
    # nobody reads or writes the padding channels as data
    for op in plan.ops:
        for v in op.reads + op.writes:
            assert v.c_off + v.c <= v.buf.c, (op.name, v.buf.name)
    if variant == "n":  # (reg_pred no longer writes a padded map: K7 decodes in its epilogue)
        # (K4: L2's bottleneck is one kernel; its expand output no longer exists as a buffer)
        assert sorted(extended) == ["L0.stem3x3s2", "L2.m0.bottleneck(k3)"]
    monkeypatch.setenv("MAFB200_PAD_FILL", "0")
    assert all(engine.pad_fill_channels(op, op.act) == 0 for op in plan.ops if op.writes)


def test_yaml_loader_accepts_reference_schema(tmp_path):
    import yaml

    rows = topology.variant_rows("s")
    path = tmp_path / "custom.yaml"
    path.write_text(yaml.safe_dump(rows))
    a, b = topology.build_graph(str(path)), topology.build_graph("s")
    assert [(l.kind, l.c_out, l.frm, l.depth, l.k) for l in a.layers] == [(l.kind, l.c_out, l.frm, l.depth, l.k) for l in b.layers]
    with pytest.raises(NotImplementedError, match="not part of the MAF-YOLO hot path"):
        bad = topology.variant_rows("n")
        bad["backbone"][2] = [-1, 1, "BotNet", [48]]
        topology.resolve(bad)


@pytest.mark.skipif(not __import__("oracle.ref_loader", fromlist=["x"]).available(), reason="reference tree not present")
def test_convert_blocks_surgery_on_reference_model():
    """convert_blocks() on a live reference Model: supported blocks are replaced in place, executor
    attributes (.i/.f) survive, Concat/Upsample/Out stay (no GPU needed until forward)."""
    import maf_yolo_b200 as mb
    from oracle import ref_loader

    m = ref_loader.build_model("n")
    before = [(mod.i, mod.f) for mod in m.backbone]
    mb.convert_blocks(m)
    kinds = [type(mod).__name__ for mod in m.backbone]
    assert kinds.count("B200Block") == 26 and kinds[11] == "Concat" and kinds[13] == "Upsample" and kinds[34] == "Out"
    assert [(mod.i, mod.f) for mod in m.backbone] == before
    assert type(m.detect).__name__ == "B200Detect" and m.detect.nc == 80
    m.float()  # reference Model._apply touches detect.stride / detect.grid (yolo.py:211-215)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 64, 64))


def test_u8_scale_is_exact_in_fp16():
    """The stem kernel turns a uint8 pixel into the fp16 operand of its MMA as fp16(u * (1/255)) — two instructions —
    where the reference computes u / 255 in fp32 (evaler.py:163) before the model's first conv sees it.  The fp32
    products differ from the quotients in the last bit for about half of the values; after the rounding to fp16 all 256
    are identical, which is what makes the shortcut bit-exact for the path."""
    import numpy as np

    u = np.arange(256, dtype=np.float32)
    quotient = u / np.float32(255)
    product = u * (np.float32(1.0) / np.float32(255))
    assert (quotient != product).sum() > 100
    assert np.array_equal(quotient.astype(np.float16), product.astype(np.float16))
