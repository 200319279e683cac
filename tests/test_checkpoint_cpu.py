"""Checkpoint ingestion (maf_yolo_b200/checkpoint.py): pickled-model `.pt` files load into a plain state_dict
without the classes they were pickled from (yolov6/utils/checkpoint.py:83-93 needs them importable)."""
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

from maf_yolo_b200 import checkpoint as ck
from oracle import ref_loader


def _fake_reference_package():
    """A throw-away package with the reference's module paths, used only to CREATE a checkpoint."""
    pkg = types.ModuleType("fakeyolo"); sub = types.ModuleType("fakeyolo.models")
    class Block(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv = nn.Conv2d(c, c, 3, bias=False)
            self.bn = nn.BatchNorm2d(c)
            self.register_buffer("scratch", torch.ones(2), persistent=False)
    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = nn.Sequential(Block(4), Block(4))
            self.detect = Block(4)
            self.detect.nc = 7
            self.yaml = {"backbone": [[-1, 1, "Block", [4]]], "nc": 7}
            self.names = ["a", "b"]
    Block.__module__ = Model.__module__ = "fakeyolo.models"
    Block.__qualname__, Model.__qualname__ = "Block", "Model"
    sub.Block, sub.Model = Block, Model
    sys.modules["fakeyolo"], sys.modules["fakeyolo.models"] = pkg, sub
    return Model


def test_loads_pickled_model_without_its_classes(tmp_path):
    Model = _fake_reference_package()
    m = Model().half()
    want = {k: v.float() for k, v in m.state_dict().items()}
    path = tmp_path / "ckpt.pt"
    torch.save({"model": m, "ema": None, "epoch": 3, "optimizer": None}, path)
    for name in ("fakeyolo", "fakeyolo.models"):
        del sys.modules[name]  # the classes are gone, as on a machine without the reference repository
    with pytest.raises(Exception):
        torch.load(path, weights_only=False)
    sd, meta = ck.load_checkpoint(path)
    assert list(sd) == list(want) and all(torch.equal(sd[k], want[k]) for k in want)
    assert all(v.dtype == torch.float32 for v in sd.values() if v.is_floating_point())
    assert "backbone.0.scratch" not in sd  # non-persistent buffers stay out, as in nn.Module.state_dict()
    assert meta["which"] == "model" and meta["yaml"]["nc"] == 7 and meta["names"] == ["a", "b"] and meta["nc"] == 7


def test_prefers_ema_and_accepts_state_dict_files(tmp_path):
    Model = _fake_reference_package()
    a, b = Model(), Model()
    torch.save({"model": a, "ema": b}, tmp_path / "e.pt")
    torch.save(a.state_dict(), tmp_path / "sd.pt")
    for name in ("fakeyolo", "fakeyolo.models"):
        del sys.modules[name]
    sd, meta = ck.load_checkpoint(tmp_path / "e.pt")
    assert meta["which"] == "ema" and torch.equal(sd["backbone.0.conv.weight"], b.backbone[0].conv.weight.detach())
    sd2, meta2 = ck.load_checkpoint(tmp_path / "sd.pt")
    assert meta2["which"] == "state_dict" and meta2["yaml"] is None
    assert torch.equal(sd2["detect.bn.running_var"], a.detect.bn.running_var)
    with pytest.raises(ValueError):
        ck.from_checkpoint(tmp_path / "sd.pt")  # a bare state_dict needs the variant


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_real_reference_checkpoint_round_trip(tmp_path):
    """A checkpoint written the way the reference's trainer writes it (yolov6/core/engine.py:195-202: the pickled
    Model object) folds to the same deploy weights as the live model."""
    from maf_yolo_b200 import fold, topology
    from oracle import model as om

    ns = ref_loader.load()
    torch.manual_seed(0)
    model = om.build_reference_model(ns, "n") if hasattr(om, "build_reference_model") else None
    if model is None:
        class NS(dict):
            __getattr__ = dict.__getitem__
        cfg = NS(model=NS(build_type="yaml", yaml_file=os.path.join(ref_loader.REF_ROOT, "configs/yaml/MAF-YOLO-n.yaml"),
                          head=NS(num_layers=3, anchors=1, strides=[8, 16, 32], use_dfl=True, reg_max=16)))
        model = ns.Model(cfg, channels=3, num_classes=80, anchors=1).eval()
    path = tmp_path / "ref.pt"
    torch.save({"model": model.half(), "ema": None, "epoch": 0}, path)
    sd, meta = ck.load_checkpoint(path)
    want = model.float().state_dict()
    assert list(sd) == list(want) and all(torch.equal(sd[k], want[k]) for k in want)
    assert meta["nc"] == 80 and meta["yaml"] is not None
    g = topology.build_graph(meta["yaml"], 80)
    a, b = fold.fold_state_dict(g, sd), fold.fold_state_dict(g, want)
    assert a.keys() == b.keys() and all(torch.equal(a[k][0], b[k][0]) and torch.equal(a[k][1], b[k][1]) for k in a)


class _Evil:
    """Pickles to REDUCE(callable, args) with a callable named by (module, name) — what a hostile checkpoint does."""

    def __init__(self, module, name, args):
        self.module, self.name, self.args = module, name, args

    def __reduce__(self):
        import importlib

        fn = importlib.import_module(self.module)
        for part in self.name.split("."):
            fn = getattr(fn, part)
        return fn, self.args


@pytest.mark.parametrize("module,name,args", [
    ("builtins", "eval", ("__import__('os').system('touch {marker}')",)),
    ("builtins", "exec", ("import os; os.system('touch {marker}')",)),
    ("os", "system", ("touch {marker}",)),
    ("posix", "system", ("touch {marker}",)),
    ("subprocess", "check_call", (["touch", "{marker}"],)),
    ("builtins", "__import__", ("os",)),
    ("builtins", "getattr", ("x", "upper")),
    ("torch.hub", "load", ("{marker}", "x")),
    ("torch.utils.cpp_extension", "load", ("x", ["{marker}"])),
    ("pathlib", "Path", ("{marker}",)),
])
def test_hostile_pickles_execute_nothing(tmp_path, module, name, args):
    """no global chosen by the file may run.  Every callable outside the explicit allow-list is an
    inert stub, so the REDUCE only records its arguments."""
    import pickle

    marker = tmp_path / "pwned"
    sub = lambda a: a.format(marker=marker) if isinstance(a, str) else ([x.format(marker=marker) for x in a] if isinstance(a, list) else a)
    payload = {"model": _Evil(module, name, tuple(sub(a) for a in args)), "ema": None}
    path = tmp_path / "evil.pt"
    with open(path, "wb") as f:  # legacy (non-zip) torch format is a plain pickle stream as well; use both
        pickle.dump(payload, f, protocol=2)
    zpath = tmp_path / "evil_zip.pt"
    torch.save(payload, zpath)
    for p in (zpath, path):
        try:
            ck.load_checkpoint(p)
        except Exception:
            pass  # "neither a pickled model checkpoint nor a state_dict" etc. is fine: nothing ran
        assert not marker.exists(), f"{module}.{name} from the checkpoint was executed"


def test_numpy_object_scalar_is_refused(tmp_path):
    """numpy's pickled `scalar(dtype('O'), bytes)` would call pickle.loads on the bytes with the STOCK unpickler."""
    import pickle
    import numpy as np

    marker = tmp_path / "pwned"
    inner = pickle.dumps(_Evil("os", "system", (f"touch {marker}",)))
    stream = (b"\x80\x02cnumpy.core.multiarray\nscalar\n" + b"cnumpy\ndtype\n(U\x02O8K\x00K\x01tR" +
              b"(K\x03U\x01|NNNJ\xff\xff\xff\xffJ\xff\xff\xff\xffK?tb" + pickle.dumps(inner, protocol=2)[2:-1] + b"\x86R.")
    path = tmp_path / "np.pt"
    path.write_bytes(stream)
    with pytest.raises(Exception):
        ck.load_checkpoint(path)
    assert not marker.exists()


def test_allow_list_is_explicit():
    up = ck._Unpickler(__import__("io").BytesIO(b""))
    for module, name in [("builtins", "eval"), ("builtins", "getattr"), ("torch.hub", "load"), ("torch", "load"),
                         ("numpy", "load"), ("torch.serialization", "load"), ("copyreg", "__reduce_ex__"),
                         ("argparse", "Namespace"), ("torch.nn.modules.conv", "Conv2d")]:
        assert issubclass(up.find_class(module, name), ck._Stub), (module, name)
    assert up.find_class("torch", "float16") is torch.float16
    assert up.find_class("collections", "OrderedDict") is __import__("collections").OrderedDict
    assert up.find_class("torch._utils", "_rebuild_tensor_v2") is torch._utils._rebuild_tensor_v2


def test_export_deploy_state_dict_round_trip(tmp_path):
    """SURVEY 8 f4, second half: folded weights -> deploy-form state_dict under the reference's keys -> folded again
    (bit-identical in fp64), and the packed weight file round trip (loaded with weights_only=True)."""
    from maf_yolo_b200 import fold, synth, topology

    for variant in ("n", "m"):
        g = topology.build_graph(variant)
        folded = fold.fold_state_dict(g, synth.random_state_dict(g, seed=2))
        sd = ck.export_deploy_state_dict(g, folded, dtype=torch.float64)
        assert "backbone.0.rbr_reparam.weight" in sd and "backbone.2.m.0.conv2.dwconv.lk_origin.bias" in sd
        assert not any(".bn." in k or "origin_bn" in k or ".norm." in k for k in sd)
        again = fold.fold_state_dict(g, sd)
        assert again.keys() == folded.keys()
        assert all(torch.equal(again[k][0], folded[k][0]) and torch.equal(again[k][1], folded[k][1]) for k in folded)
    path = tmp_path / "n.mafb200"
    ck.save_packed(path, "n", folded if variant == "n" else fold.fold_state_dict(topology.build_graph("n"), synth.random_state_dict(topology.build_graph("n"), seed=2)),
                   names=[f"c{i}" for i in range(80)])
    f2, rows, nc, names = ck.load_packed(path)
    assert nc == 80 and names[3] == "c3" and rows["backbone"][0][2] == "RepVGGBlock"
    gn = topology.build_graph(rows, nc)
    assert [l.kind for l in gn.layers] == [l.kind for l in topology.build_graph("n").layers]
    ref = fold.fold_state_dict(gn, synth.random_state_dict(gn, seed=2))
    assert all(torch.equal(f2[k][0], ref[k][0].half().double()) and torch.equal(f2[k][1], ref[k][1].float().double()) for k in ref)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_exported_state_dict_loads_into_the_reference_deploy_model():
    """The exported state_dict loads strictly into the reference's own deploy-form model, whose forward then equals
    the train-form forward of the original weights (the reference does the same folding in fp32)."""
    from maf_yolo_b200 import fold, synth, topology

    g = topology.build_graph("n")
    sd_train = synth.random_state_dict(g, seed=4)
    m = ref_loader.build_model("n")
    m.load_state_dict(sd_train, strict=True)
    x = torch.rand(1, 3, 320, 320, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        want = m(x)[0]
    dep = ref_loader.to_deploy(ref_loader.build_model("n"))
    exported = ck.export_deploy_state_dict(g, fold.fold_state_dict(g, sd_train))
    missing = dep.load_state_dict(exported, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        got = dep(x)[0]
    assert (got[..., :4] - want[..., :4]).abs().max() < 2e-3 and (got[..., 5:] - want[..., 5:]).abs().max() < 1e-5
