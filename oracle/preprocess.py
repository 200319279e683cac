"""CPU restatement of the reference's image pre-processing (TEST INFRASTRUCTURE — never imported by the
product package).  Pure numpy, so it also runs on the GPU box where neither the reference nor its cv2 are
needed:

  letterbox_geometry / letterbox   yolov6/data/data_augment.py:53-83
  resize_linear_u8                 cv2.resize(..., interpolation=cv2.INTER_LINEAR) for uint8 — OpenCV
                                   modules/imgproc/src/resize.cpp (third-party; opencv-python 4.13.0 in this image):
                                   11-bit fixed-point coefficients, x weights zeroed at the borders, y source
                                   rows clipped with their weights kept, VResizeLinear's shifted sum
  precess_image                    yolov6/core/inferer.py:168-178  (HWC -> CHW, BGR -> RGB, /255)

Pinned bit-exactly against cv2.resize and the reference's letterbox / Inferer.precess_image in
tests/test_preprocess_cpu.py (build container) and against tests/golden/preprocess.npz.
"""
from __future__ import annotations

import numpy as np


def letterbox_geometry(shape, new_shape=(640, 640), auto=True, scaleup=True, stride=32):
    """(h, w) of the source -> dict(r, new_unpad=(w,h), dw, dh, top, bottom, left, right) as data_augment.py:55-76."""
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


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw, _ = src.shape

    def coeffs(dn, sn, zero_at_border):
        scale = sn / dn
        d = np.arange(dn, dtype=np.float64)
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if zero_at_border:
            lo = s < 0
            f[lo] = 0
            s[lo] = 0
            hi = s >= sn - 1
            f[hi] = 0
            s[hi] = sn - 1
        a0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int64)
        a1 = np.rint(f * np.float32(2048)).astype(np.int64)
        return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1

    sx, sx1, ax0, ax1 = coeffs(dw, sw, True)
    sy, sy1, by0, by1 = coeffs(dh, sh, False)
    S = src.astype(np.int64)
    H = S[:, sx, :] * ax0[None, :, None] + S[:, sx1, :] * ax1[None, :, None]
    r0, r1 = H[sy], H[sy1]
    out = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(im: np.ndarray, new_shape=(640, 640), color=(114, 114, 114), auto=True, scaleup=True, stride=32,
              return_int=False):
    g = letterbox_geometry(im.shape[:2], new_shape, auto, scaleup, stride)
    if im.shape[:2][::-1] != g["new_unpad"]:
        im = resize_linear_u8(im, g["new_unpad"][0], g["new_unpad"][1])
    h, w = im.shape[:2]
    out = np.empty((h + g["top"] + g["bottom"], w + g["left"] + g["right"], 3), dtype=np.uint8)
    out[...] = np.asarray(color, dtype=np.uint8)
    out[g["top"]:g["top"] + h, g["left"]:g["left"] + w] = im
    if not return_int:
        return out, g["r"], (g["dw"], g["dh"])
    return out, g["r"], (g["left"], g["top"])


def precess_image(img_src: np.ndarray, img_size, stride, half=False):
    image = letterbox(img_src, img_size, stride=stride)[0]
    image = np.ascontiguousarray(image.transpose((2, 0, 1))[::-1])
    return image, (image.astype(np.float16 if half else np.float32) / 255)
