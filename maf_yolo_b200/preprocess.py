"""Image pre-processing of the reference on the device, behind the reference's own names.

  letterbox(im, new_shape, color, auto, scaleup, stride, return_int)   == yolov6/data/data_augment.py:53-83
  precess_image(img_src, img_size, stride, half)                        == Inferer.precess_image (inferer.py:168-178)
  letterbox_batch(images, new_shape, orig_shapes=...)                   the evaluator's letterbox step for a fixed-size
                                                                        batch (datasets.py:204-215, auto=False); input =
                                                                        the output of load_image (datasets.py:280-301)

The image is uploaded once as the raw uint8 HWC (BGR) array cv2 produced; ONE kernel (mafb200_letterbox_u8)
resizes with OpenCV's exact 8-bit INTER_LINEAR arithmetic, pads with 114, converts HWC -> CHW / BGR -> RGB and
writes the uint8 tensor the model consumes directly (its stem kernel folds the /255), so the fp32 image
(4x the bytes) never exists.  Geometry is computed on the host with the reference's own expressions.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib


def letterbox_geometry(shape, new_shape=(640, 640), auto=True, scaleup=True, stride=32):
    """Same arithmetic as data_augment.py:55-76 (python floats / numpy mod), returned as a dict."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return dict(r=r, new_unpad=new_unpad, dw=dw, dh=dh, top=top, bottom=bottom, left=left, right=right)


def _to_device_u8(im, device) -> torch.Tensor:
    if isinstance(im, np.ndarray):
        if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
            raise TypeError("expected an HWC uint8 image with 3 channels")
        im = torch.from_numpy(np.ascontiguousarray(im))
    if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
        raise TypeError("expected an HWC uint8 image with 3 channels")
    return im.to(device, non_blocking=True).contiguous()


def _launch(src: torch.Tensor, dst: torch.Tensor, g: dict, color, swap_rb: bool) -> None:
    if len(set(int(c) for c in color)) != 1:
        raise NotImplementedError("only a grey border (the reference's 114,114,114) is supported")
    h0, w0, _ = src.shape
    check(lib().mafb200_letterbox_u8(src.data_ptr(), h0, w0, 3 * w0, dst.data_ptr(), dst.shape[1], dst.shape[2],
                                     g["new_unpad"][1], g["new_unpad"][0], g["top"], g["left"], int(color[0]),
                                     int(swap_rb), torch.cuda.current_stream().cuda_stream))


def letterbox(im, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleup=True, stride=32, return_int=False,
              device="cuda", swap_rb: bool = False):
    """Reference signature; returns (image, r, (dw, dh)) with image a CUDA uint8 **CHW** tensor (channel order
    kept unless swap_rb) — `image.permute(1, 2, 0)` is the reference's HWC array."""
    if not torch.cuda.is_available():
        raise RuntimeError("maf_yolo_b200.preprocess needs a B200 GPU (no CPU fallback)")
    src = _to_device_u8(im, device)
    g = letterbox_geometry(src.shape[:2], new_shape, auto, scaleup, stride)
    out_h = g["new_unpad"][1] + g["top"] + g["bottom"]
    out_w = g["new_unpad"][0] + g["left"] + g["right"]
    with torch.cuda.device(src.device):
        dst = torch.empty((3, out_h, out_w), dtype=torch.uint8, device=src.device)
        _launch(src, dst, g, color, swap_rb)
    if not return_int:
        return dst, g["r"], (g["dw"], g["dh"])
    return dst, g["r"], (g["left"], g["top"])


def precess_image(img_src, img_size, stride, half=False, device="cuda", as_uint8: bool = True):
    """Inferer.precess_image (sic): letterbox + HWC->CHW + BGR->RGB.  Returns (image, img_src); image is the CUDA
    uint8 [3,H,W] tensor the B200 model takes directly (as_uint8=True, default) or the reference's float/half
    tensor in [0,1] (as_uint8=False)."""
    image = letterbox(img_src, img_size, stride=stride, device=device, swap_rb=True)[0]
    if not as_uint8:
        image = image.half() if half else image.float()
        image /= 255
    return image, img_src


def letterbox_batch(images: Sequence, new_shape=(640, 640), scaleup: bool = False, device="cuda",
                    color=(114, 114, 114), orig_shapes: Sequence[Tuple[int, int]] = None) -> Tuple[torch.Tensor, List[tuple]]:
    """The evaluator's letterbox step (datasets.py:204-215: `letterbox(img, shape, auto=False, scaleup=self.augment)`
    with augment=False) for a fixed-size batch: uint8 [B,3,H,W] RGB on the device + per image the dataloader's
    `shapes` = ((h0, w0), ((h*ratio/h0, w*ratio/w0), pad)) (datasets.py:215) for the post-NMS rescale.

    Contract: `images[i]` is what the reference's `load_image` RETURNS (datasets.py:280-301) — the file already
    resized so that its long side equals img_size (cv2 INTER_AREA when shrinking, INTER_LINEAR when enlarging,
    `int()` geometry) — NOT the original file.  That resize is the step before this one and is not reproduced here
    (its INTER_AREA path is a different algorithm); pass the original sizes as `orig_shapes[i]` = (h0, w0) so that
    `shapes` carries the reference's values.  Without `orig_shapes` the images are taken to be the originals of a
    dataset whose long side already equals img_size (h0, w0 = h, w).  With scaleup=False a smaller image is only
    padded, exactly as the reference's letterbox does to it."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    if not torch.cuda.is_available():
        raise RuntimeError("maf_yolo_b200.preprocess needs a B200 GPU (no CPU fallback)")
    if orig_shapes is not None and len(orig_shapes) != len(images):
        raise ValueError("orig_shapes must give (h0, w0) for every image")
    dev = torch.device(device)
    with torch.cuda.device(dev):
        batch = torch.empty((len(images), 3, new_shape[0], new_shape[1]), dtype=torch.uint8, device=dev)
        shapes = []
        for i, im in enumerate(images):
            src = _to_device_u8(im, dev)
            h, w = src.shape[:2]
            h0, w0 = (h, w) if orig_shapes is None else (int(orig_shapes[i][0]), int(orig_shapes[i][1]))
            g = letterbox_geometry((h, w), new_shape, False, scaleup, 32)
            _launch(src, batch[i], g, color, True)
            shapes.append(((h0, w0), ((h * g["r"] / h0, w * g["r"] / w0), (g["dw"], g["dh"]))))
    return batch, shapes


def load_image_size(h0: int, w0: int, img_size: int) -> Tuple[int, int]:
    """(h, w) the reference's load_image resizes an (h0, w0) file to (datasets.py:290-300: r = img_size / max,
    `int(w0 * r), int(h0 * r)`), for callers that do that resize themselves."""
    r = img_size / max(h0, w0)
    return (h0, w0) if r == 1 else (int(h0 * r), int(w0 * r))
