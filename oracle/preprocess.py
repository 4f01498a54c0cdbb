"""CPU restatement of the reference's image input pipeline (the ORACLE for SURVEY 8f-3).  TEST INFRASTRUCTURE ONLY.

Follows data/dataset.py:62-83 (LoadDataset.__getitem__: PIL image -> optional 90-degree rotation of tall images ->
numpy -> transform) and data/augmentations.py:27-46 (get_valid_transforms / get_test_transforms):
``A.Resize(height, width)`` = ``cv2.resize(img, (width, height), interpolation=cv2.INTER_LINEAR)`` on the uint8 image,
``A.Normalize(mean, std)`` = ``(img - mean*255) * (1 / (std*255))`` in float32, ``ToTensorV2`` = HWC -> CHW.

albumentations is a third-party dependency absent from /root/reference and from this image (requirements.txt pins
albumentations==1.0.0); its two transforms are restated from their published definition.  ``cv2`` IS in the image, so
the resize -- the only non-trivial arithmetic -- is pinned bit-exactly against cv2.resize itself
(tests/test_oracle_preprocess.py): OpenCV's 8-bit INTER_LINEAR works in fixed point (11-bit coefficients, the
``((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2`` vertical pass) and switches to a 2x2 box average when both scale
factors are exactly 2 (imgproc/src/resize.cpp).
"""
from __future__ import annotations

import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _coeffs(src: int, dst: int):
    inv = float(dst) / float(src)
    scale = 1.0 / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    return s, f - s.astype(np.float32)


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR) for uint8 [H, W] or [H, W, C], bit for bit."""
    h, w = img.shape[:2]
    if h == 2 * out_h and w == 2 * out_w:      # resize.cpp: INTER_LINEAR with both scales exactly 2 becomes INTER_AREA
        im = img.astype(np.int64)
        return ((im[0::2, 0::2] + im[0::2, 1::2] + im[1::2, 0::2] + im[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    im = img.reshape(h, w, -1).astype(np.int64)
    sx, fx = _coeffs(w, out_w)
    lo = sx < 0
    fx, sx = np.where(lo, np.float32(0), fx), np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx, sx = np.where(hi, np.float32(0), fx), np.where(hi, w - 1, sx)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    hor = im[:, sx, :] * a0[None, :, None] + im[:, np.minimum(sx + 1, w - 1), :] * a1[None, :, None]
    sy, fy = _coeffs(h, out_h)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    s0, s1 = hor[np.clip(sy, 0, h - 1)], hor[np.clip(sy + 1, 0, h - 1)]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((out_h, out_w) + img.shape[2:])


def normalize(img_u8: np.ndarray) -> np.ndarray:
    """A.Normalize(mean, std, max_pixel_value=255) -> float32 [H, W, C].  A single-channel image (data.rgb = 1) takes the
    first channel's statistics (the 3-element mean does not broadcast against an [H, W] array in the reference)."""
    c = 1 if img_u8.ndim == 2 else img_u8.shape[2]
    mean = np.array(MEAN[:c], np.float32) * np.float32(255)
    den = np.reciprocal(np.array(STD[:c], np.float32) * np.float32(255), dtype=np.float32)
    x = img_u8.reshape(img_u8.shape[0], img_u8.shape[1], c).astype(np.float32)
    x -= mean
    x *= den
    return x


def load_item(img_u8: np.ndarray, height: int, width: int) -> np.ndarray:
    """dataset.py:75-81 + get_valid_transforms: uint8 [H, W] or [H, W, C] -> float32 [C, height, width]."""
    h, w = img_u8.shape[:2]
    if h / w > 2:                                # dataset.py:77-78  image.rotate(90, expand=True)
        img_u8 = np.rot90(img_u8, 1)
    x = normalize(resize_linear_u8(np.ascontiguousarray(img_u8), height, width))
    return np.ascontiguousarray(x.transpose(2, 0, 1))
