"""maf_yolo_b200 — B200-native (sm_100a) forward/detect hot path of MAF-YOLO behind a C ABI.

Public API (mirrors the reference's names):
    convert(model)                    reference `Model`  -> B200DetectModel        (nn.py)
    from_state_dict(sd, variant)      reference weights  -> B200DetectModel
    non_max_suppression(...)          == yolov6/utils/nms.py:31
    from_checkpoint(path)             reference `.pt` (pickled model) -> B200DetectModel, no reference code needed (checkpoint.py)
    export_deploy_state_dict(...)     folded weights -> deploy-form state_dict the reference loads; save_packed / from_packed
    preprocess.* / postprocess.*      letterbox / precess_image and scale_coords / convert_to_coco_format on the device
    convert_blocks(model)             per-block drop-ins inside the reference's own Model.forward loop (blocks.py)
"""
from .checkpoint import (export_deploy_state_dict, from_checkpoint, from_packed, load_checkpoint, load_packed,  # noqa: F401
                         save_packed)
from .blocks import B200Block, B200Detect, convert_blocks  # noqa: F401
from .nn import B200DetectModel, DetectTicket, convert, from_state_dict, non_max_suppression, non_max_suppression_padded  # noqa: F401

__all__ = ["B200DetectModel", "DetectTicket", "convert", "from_state_dict", "non_max_suppression", "non_max_suppression_padded",
           "B200Block", "B200Detect", "convert_blocks", "from_checkpoint", "load_checkpoint", "export_deploy_state_dict",
           "save_packed", "load_packed", "from_packed"]
