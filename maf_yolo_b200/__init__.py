"""maf_yolo_b200 — B200-native (sm_100a) forward/detect hot path of MAF-YOLO behind a C ABI.

Public API (mirrors the reference's names):
    convert(model)                    reference `Model`  -> B200DetectModel        (nn.py)
    from_state_dict(sd, variant)      reference weights  -> B200DetectModel
    non_max_suppression(...)          == yolov6/utils/nms.py:31
    convert_blocks(model)             per-block drop-ins inside the reference's own Model.forward loop (blocks.py)
"""
from .blocks import B200Block, B200Detect, convert_blocks  # noqa: F401
from .nn import B200DetectModel, DetectTicket, convert, from_state_dict, non_max_suppression, non_max_suppression_padded  # noqa: F401

__all__ = ["B200DetectModel", "DetectTicket", "convert", "from_state_dict", "non_max_suppression", "non_max_suppression_padded",
           "B200Block", "B200Detect", "convert_blocks"]
