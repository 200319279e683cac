"""Checkpoint ingestion: the reference's pickled `.pt` files without the reference's classes.

`yolov6/utils/checkpoint.py:83-93` (`load_checkpoint`) does `torch.load(weights)` and takes `ckpt['ema']` or
`ckpt['model']` — a pickled *nn.Module object*, so unpickling normally needs yolov6.models.yolo.Model,
yolov6.layers.common.* ... importable under exactly those paths.  Here a restricted unpickler resolves only
an explicit allow-list of tensor-rebuild functions and plain containers; every other global (the reference's modules, its Config objects, and any
torch / numpy / builtins name that is not on the explicit allow-list below, e.g. builtins.eval or torch.hub.load)
becomes an inert stub that just holds its pickled `__dict__`.  The module tree is then walked (`_modules`,
`_parameters`, `_buffers`) into an ordinary `state_dict`, and `model.yaml` (yolo.py:145), `names` and `nc` are
read off the stub — everything `from_state_dict` / `convert` need.  No code from the checkpoint is executed.

    model = maf_yolo_b200.from_checkpoint("MAFYOLOn.pt")          # -> B200DetectModel
    sd, meta = maf_yolo_b200.checkpoint.load_checkpoint("x.pt")   # state_dict (fp32) + {"yaml", "names", "nc", "which"}
"""
from __future__ import annotations

import pickle
import types
from collections import OrderedDict
from typing import Dict, Tuple

import torch

# Explicit allow-list of (module, name) globals the unpickler may resolve to the REAL object: exactly what
# torch.save needs to rebuild tensors / parameters and plain containers.  Everything else — the reference's classes,
# torch.nn modules, argparse / pathlib objects, and every other torch / numpy / builtins name (eval, exec, getattr,
# __import__, torch.hub.load, torch.utils.cpp_extension.load, numpy's pickle-calling `scalar`, ...) — becomes an inert
# stub, so no callable chosen by the file ever runs.
_TORCH_STORAGES = ("DoubleStorage", "FloatStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage",
                   "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "ComplexFloatStorage",
                   "ComplexDoubleStorage", "UntypedStorage")
_ALLOWED = {
    ("collections", "OrderedDict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"),
    ("torch._utils", "_rebuild_parameter"), ("torch._utils", "_rebuild_parameter_with_state"),
    ("torch._tensor", "_rebuild_from_type_v2"),
    ("torch.nn.parameter", "Parameter"), ("torch", "Tensor"), ("torch", "Size"), ("torch", "device"),
    ("torch.storage", "UntypedStorage"), ("torch.storage", "TypedStorage"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("_codecs", "encode"),
} | {("torch", n) for n in _TORCH_STORAGES} | {
    (m, n) for m in ("builtins", "__builtin__")
    for n in ("set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "bytearray", "slice",
              "complex", "range", "object")}
_stub_cache: Dict[Tuple[str, str], type] = {}


class _Stub:
    """Stands in for any global of the checkpoint that is not on the allow-list: keeps the pickled state only.
    Calling it, constructing it, or filling it (SETITEM / APPEND opcodes of dict / list subclasses) has no effect
    beyond storing the arguments."""

    def __init__(self, *args, **kwargs):
        self._stub_args = (args, kwargs)

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2 and isinstance(state[0], (dict, type(None))):
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)
        else:
            self.__dict__["_stub_state"] = state

    def __call__(self, *args, **kwargs):  # e.g. a pickled functools.partial-like reduce on a stubbed callable
        return _Stub(*args, **kwargs)

    def __setitem__(self, key, value):
        self.__dict__.setdefault("_stub_items", {})[key] = value

    def append(self, value):
        self.__dict__.setdefault("_stub_list", []).append(value)

    def extend(self, values):
        self.__dict__.setdefault("_stub_list", []).extend(values)

    def add(self, value):
        self.__dict__.setdefault("_stub_list", []).append(value)


def _stub_class(module: str, name: str) -> type:
    key = (module, name)
    cls = _stub_cache.get(key)
    if cls is None:
        cls = type(name.rsplit(".", 1)[-1], (_Stub,), {"__module__": module})
        _stub_cache[key] = cls
    return cls


def _safe_reconstructor(cls, base, state):
    """copyreg._reconstructor for protocol < 2 pickles of plain objects: only ever instantiates stubs."""
    if isinstance(cls, type) and issubclass(cls, _Stub):
        obj = object.__new__(cls)
        if state is not None:
            obj.__dict__["_stub_base_state"] = state
        return obj
    raise pickle.UnpicklingError(f"refusing to reconstruct {cls!r} from a checkpoint")


def _safe_numpy_scalar(dtype, data):
    """numpy's pickled scalars; the real `scalar` unpickles `data` again when dtype is object — never allowed."""
    import numpy as np

    if not isinstance(dtype, np.dtype) or dtype.hasobject:
        raise pickle.UnpicklingError("object-dtype numpy scalar in a checkpoint")
    return np.frombuffer(data, dtype=dtype, count=1)[0]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "copyreg" and name == "_reconstructor":
            return _safe_reconstructor
        if module in ("numpy.core.multiarray", "numpy._core.multiarray") and name == "scalar":
            return _safe_numpy_scalar
        if (module, name) in _ALLOWED:
            return super().find_class(module, name)
        if module == "torch" and isinstance(getattr(torch, name, None), torch.dtype):
            return getattr(torch, name)
        return _stub_class(module, name)


# the object torch.load(pickle_module=...) expects: a module-like namespace with Unpickler / load
_pickle_module = types.ModuleType("maf_yolo_b200._restricted_pickle")
_pickle_module.Unpickler = _Unpickler
_pickle_module.load = lambda f, **kw: _Unpickler(f, **kw).load()
_pickle_module.__dict__.update({k: getattr(pickle, k) for k in ("HIGHEST_PROTOCOL", "DEFAULT_PROTOCOL", "PickleError",
                                                                 "UnpicklingError", "dump", "dumps", "Pickler")})


def _is_module_like(obj) -> bool:
    d = getattr(obj, "__dict__", None)
    return isinstance(d, dict) and "_modules" in d and "_parameters" in d


def module_state_dict(mod, prefix: str = "", out=None) -> "OrderedDict[str, torch.Tensor]":
    """state_dict() of a (stubbed or real) module tree, same keys and order as nn.Module.state_dict."""
    out = OrderedDict() if out is None else out
    d = mod.__dict__
    for k, p in (d.get("_parameters") or {}).items():
        if p is not None:
            out[prefix + k] = p.detach() if isinstance(p, torch.Tensor) else p
    skip = d.get("_non_persistent_buffers_set") or set()
    for k, b in (d.get("_buffers") or {}).items():
        if b is not None and k not in skip:
            out[prefix + k] = b
    for k, m in (d.get("_modules") or {}).items():
        if m is not None:
            module_state_dict(m, prefix + k + ".", out)
    return out


def load_checkpoint(weights, map_location="cpu"):
    """Same selection rule as the reference (`ckpt['ema'] if ckpt.get('ema') else ckpt['model']`, then `.float()`).
    Also accepts a bare pickled module, or a plain (train- or deploy-form) state_dict file.
    Returns (state_dict, meta) with meta = {"yaml": model.yaml rows | None, "names", "nc", "which"}."""
    # weights_only=False only because torch.load rejects a custom pickle_module otherwise; the pickle_module IS the
    # restriction (explicit allow-list above; tests/test_checkpoint_cpu.py loads hostile pickles through it)
    ckpt = torch.load(weights, map_location=map_location, pickle_module=_pickle_module, weights_only=False)
    which, model = "state_dict", None
    if isinstance(ckpt, dict) and ("model" in ckpt or "ema" in ckpt):
        which = "ema" if ckpt.get("ema") is not None and ckpt.get("ema") is not False else "model"
        model = ckpt[which]
    elif _is_module_like(ckpt):
        which, model = "module", ckpt
    if model is not None and _is_module_like(model):
        sd = module_state_dict(model)
        meta = {"yaml": getattr(model, "yaml", None), "names": getattr(model, "names", None), "which": which}
        det = (model.__dict__.get("_modules") or {}).get("detect")
        meta["nc"] = getattr(det, "nc", None) if det is not None else None
    else:
        sd = model if model is not None else ckpt
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        if not (isinstance(sd, dict) and sd and all(isinstance(v, torch.Tensor) for v in sd.values())):
            raise TypeError(f"{weights}: neither a pickled model checkpoint nor a state_dict")
        meta = {"yaml": None, "names": None, "nc": None, "which": which}
    sd = OrderedDict((k, v.float() if v.is_floating_point() else v) for k, v in sd.items())
    return sd, meta


def from_checkpoint(weights, variant_or_yaml=None, nc=None, bn_eps: float = 1e-3, **kw):
    """Reference `.pt` -> B200DetectModel.  The topology comes from the checkpoint's own `model.yaml` unless
    `variant_or_yaml` ('n' | 's' | 'm' | yaml path / rows) is given (needed for bare state_dict files)."""
    from .nn import from_state_dict

    sd, meta = load_checkpoint(weights)
    topo = variant_or_yaml if variant_or_yaml is not None else meta["yaml"]
    if topo is None:
        raise ValueError("the file holds a bare state_dict: pass variant_or_yaml ('n' | 's' | 'm' | yaml)")
    ncls = nc if nc is not None else (meta["nc"] or 80)
    if meta.get("names") is not None and "names" not in kw:
        kw["names"] = meta["names"]
    return from_state_dict(sd, topo, ncls, bn_eps, **kw)


# ------------------------------------------------------------------------------------------------
# export (second half): folded weights back to the reference / to a packed file
# ------------------------------------------------------------------------------------------------
def export_deploy_state_dict(graph, folded, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """The folded (weight, bias) pairs as a DEPLOY-form `state_dict` under the reference's own keys — what its model
    holds after `fuse_model` + `switch_to_deploy` + `reparameterize` (yolov6/core/evaler.py:93-109): `rbr_reparam.*` for
    RepVGGBlocks, `conv.weight` / `conv.bias` for every fused `Conv`, `dwconv.lk_origin.*` for the merged depth-wise
    kernels, `cls_pred` / `reg_pred`, and the DFL projection.  It loads into a reference deploy model with
    `load_state_dict(strict=True)` and into this package again (`from_state_dict` detects the form by key presence)."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def put(prefix, key, names=("weight", "bias")):
        w, b = folded[key]
        sd[f"{prefix}.{names[0]}"] = w.to(dtype).contiguous()
        sd[f"{prefix}.{names[1]}"] = b.to(dtype).contiguous()

    for l in graph.layers:
        p, i = f"backbone.{l.i}", str(l.i)
        if l.kind == "repvgg":
            put(p + ".rbr_reparam", i)
        elif l.kind == "rephdw":
            put(p + ".conv1.conv", i + ".conv1")
            for j in range(l.depth):
                put(f"{p}.m.{j}.conv1.conv", f"{i}.m.{j}.conv1")
                put(f"{p}.m.{j}.conv2.dwconv.lk_origin", f"{i}.m.{j}.dw")
                put(f"{p}.m.{j}.one_conv.conv", f"{i}.m.{j}.one_conv")
            put(p + ".conv2.conv", i + ".conv2")
        elif l.kind == "mprep":
            put(p + ".conv1.conv", i + ".conv1")
            put(p + ".conv2.rbr_reparam", i + ".conv2")
        elif l.kind == "sppf":
            put(p + ".cv1.conv", i + ".cv1")
            put(p + ".cv2.conv", i + ".cv2")
        elif l.kind == "convw":
            put(p + ".block.conv", i + ".block")
        elif l.kind == "head":
            put(p + ".stem.conv", i + ".stem")
            for br in ("cls", "reg"):
                put(f"{p}.{br}_conv.dwconv.lk_origin", f"{i}.{br}_dw")
                put(f"{p}.{br}_conv_s.conv", f"{i}.{br}_s")
                put(f"{p}.{br}_pred", f"{i}.{br}_pred")
    reg_max = graph.layers[graph.head_layers[0]].reg_max
    sd["detect.proj"] = torch.linspace(0, reg_max, reg_max + 1).to(dtype)
    sd["detect.proj_conv.weight"] = sd["detect.proj"].view(1, reg_max + 1, 1, 1).clone()
    return sd


def save_packed(path, variant_or_yaml, folded, names=None, nc: int = 80, dtype=torch.float16) -> None:
    """Packed weight file: the folded deploy-form tensors (fp16 by default — released checkpoints carry fp16 precision,
    yolov6/utils/checkpoint.py:119) keyed as `fold_state_dict` keys them + the topology.  Plain tensors / containers
    only: `load_packed` reads it with `torch.load(weights_only=True)`."""
    from .topology import variant_rows

    rows = variant_rows(variant_or_yaml) if isinstance(variant_or_yaml, str) and variant_or_yaml in ("n", "s", "m") else variant_or_yaml
    blob = {"format": "mafb200-packed-1", "yaml": rows, "nc": int(nc), "names": list(names) if names is not None else None,
            "weights": {k: (w.to(dtype).contiguous(), b.to(torch.float32).contiguous()) for k, (w, b) in folded.items()}}
    torch.save(blob, path)


def load_packed(path):
    """-> (folded, yaml rows, nc, names) of a `save_packed` file."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if not isinstance(blob, dict) or blob.get("format") != "mafb200-packed-1":
        raise TypeError(f"{path}: not a mafb200 packed weight file")
    folded = {k: (w.to(torch.float64), b.to(torch.float64)) for k, (w, b) in blob["weights"].items()}
    return folded, blob["yaml"], blob["nc"], blob["names"]


def from_packed(path, **kw):
    """Packed weight file -> B200DetectModel (no folding at load time)."""
    from .nn import B200DetectModel
    from .topology import build_graph

    folded, rows, nc, names = load_packed(path)
    return B200DetectModel(build_graph(rows, nc), folded, names=names, **kw)
