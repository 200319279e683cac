"""ctypes binding of libmafb200.so (the C ABI declared in include/mafb200.h).

The library is the product: there is no Python/PyTorch fallback for any op.  If the shared
object is missing or a call fails, this module raises — loudly.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# MAFB200_LIB: an A/B build of the SAME library (maf_yolo_b200/build.py, MAFB200_BUILD_SUFFIX) — never a fallback
LIB_PATH = Path(os.environ.get("MAFB200_LIB") or Path(__file__).resolve().parent / "libmafb200.so")

MAF_F16, MAF_F32, MAF_U8 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
ACT_CODES = {"none": ACT_NONE, "silu": ACT_SILU, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID}
MAX_SRC = 4


class MafTensor(C.Structure):
    """Mirror of `maf_tensor` (include/mafb200.h): an NHWC fp16 view, possibly a channel slice."""

    _fields_ = [
        ("ptr", C.c_void_p),
        ("n", C.c_int32),
        ("h", C.c_int32),
        ("w", C.c_int32),
        ("c", C.c_int32),
        ("c_stride", C.c_int32),
        ("dtype", C.c_int32),
    ]

    def __repr__(self):
        return f"MafTensor(ptr=0x{self.ptr or 0:x}, n={self.n}, h={self.h}, w={self.w}, c={self.c}, ld={self.c_stride})"


class MafError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libmafb200 error {code}: {text}")
        self.code = code


_P = C.POINTER
_SIGNATURES = {
    "mafb200_version": (C.c_int32, []),
    "mafb200_last_error": (C.c_char_p, []),
    "mafb200_device_ok": (C.c_int32, [C.c_int32]),
    "mafb200_gemm_tiling": (C.c_int32, [C.c_int32, _P(C.c_int32), _P(C.c_int32)]),
    "mafb200_packed_k_1x1": (C.c_int32, [_P(C.c_int32), C.c_int32]),
    "mafb200_packed_k_3x3": (C.c_int32, [C.c_int32]),
    "mafb200_conv1x1": (C.c_int32, [_P(MafTensor), C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, _P(MafTensor),
                                    _P(MafTensor), C.c_void_p]),
    "mafb200_conv3x3s2": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_conv3x3s2_pair": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_stem_conv3x3s2": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_dwconv": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, _P(MafTensor),
                                   C.c_void_p]),
    "mafb200_dwconv_conv1x1": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_bottleneck_supported": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "mafb200_bottleneck_trace": (C.c_int32, [C.c_void_p]),
    "mafb200_bottleneck": (C.c_int32, [_P(MafTensor), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, _P(MafTensor), C.c_void_p]),
    "mafb200_maxpool2x2_conv1x1": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_maxpool2x2": (C.c_int32, [_P(MafTensor), _P(MafTensor), C.c_void_p]),
    "mafb200_sppf_pool": (C.c_int32, [_P(MafTensor), _P(MafTensor), _P(MafTensor), _P(MafTensor), C.c_void_p]),
    "mafb200_upsample2x": (C.c_int32, [_P(MafTensor), _P(MafTensor), C.c_void_p]),
    "mafb200_nchw_to_nhwc_f16": (C.c_int32, [C.c_void_p, C.c_int32, _P(MafTensor), C.c_void_p]),
    "mafb200_nhwc_f16_to_nchw": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_head_decode": (C.c_int32, [_P(MafTensor), _P(MafTensor), _P(C.c_float), C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p]),
    "mafb200_nms_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "mafb200_nms": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p]),
    "mafb200_letterbox_u8": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "mafb200_scale_detections": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mafb200_head_decode_detect": (C.c_int32, [_P(MafTensor), _P(MafTensor), _P(C.c_float), C.c_int32, C.c_int32, C.c_int32,
                                               C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.c_void_p]),
    "mafb200_nms_select": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mafb200_detect_cfg_fill": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]),
    "mafb200_detect_reset": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_head_pred": (C.c_int32, [_P(MafTensor), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mafb200_nms_select_packed": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                              C.c_void_p]),
    "mafb200_loss_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "mafb200_detect_loss": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mafb200_launch_count": (C.c_int64, []),
}
DETECT_CFG_BYTES = 16 + 256  # sizeof(maf_detect_cfg)
HEAD_CLS, HEAD_REG, HEAD_CLS_TRAIN, HEAD_REG_TRAIN = 0, 1, 2, 3

_lib = None


def exported_symbols() -> list[str]:
    """Every entry point include/mafb200.h declares (the CPU test checks the .so exports them all)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    """Loads libmafb200.so once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m maf_yolo_b200.build` "
                "(or __graft_entry__.build()).  There is no CPU/PyTorch fallback for this path.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MafError(rc, lib().mafb200_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().mafb200_launch_count())


def gemm_tiling(cout: int) -> tuple[int, int]:
    nt, tn = C.c_int32(), C.c_int32()
    check(lib().mafb200_gemm_tiling(cout, C.byref(nt), C.byref(tn)))
    return nt.value, tn.value


def packed_k_1x1(src_channels) -> int:
    arr = (C.c_int32 * len(src_channels))(*src_channels)
    k = lib().mafb200_packed_k_1x1(arr, len(src_channels))
    if k < 0:
        check(k)
    return k


def packed_k_3x3(cin: int) -> int:
    k = lib().mafb200_packed_k_3x3(cin)
    if k < 0:
        check(k)
    return k
