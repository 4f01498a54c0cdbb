"""Image input pipeline on the device (data/dataset.py:62-83 + data/augmentations.py:27-46, valid / test transforms).

``transform_batch`` takes the decoded uint8 images of a mini-batch (what ``np.array(PIL.Image)`` gives the reference's
transform: [H, W] for ``data.rgb = 1``, [H, W, 3] for 3) and returns the collated fp32 tensor [B, C, height, width] on
the GPU: the rotation of tall images, cv2's INTER_LINEAR resize (bit-exact 8-bit fixed point), Normalize and the
HWC -> CHW transpose run in one kernel; only the raw uint8 pixels cross PCIe (4x fewer bytes than the fp32 tensors the
reference's DataLoader ships, 12x fewer for images larger than the network input).  Training-time augmentation
(ShiftScaleRotate / GridDistortion, augmentations.py:5-24) draws from albumentations' RNG and stays on the host."""
import ctypes

import numpy as np
import torch

from . import _lib

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def transform_batch(images, height, width, device="cuda", rotate_tall=True):
    """images: sequence of uint8 numpy arrays / CPU tensors, all with the same channel count -> fp32 [B, C, height, width]."""
    lib = _lib.load_library()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("frx runs on a CUDA device (B200) only; there is no CPU fallback")
    arrs = [np.ascontiguousarray(np.asarray(im)) for im in images]
    if not arrs:
        raise ValueError("empty batch")
    ch = 1 if arrs[0].ndim == 2 else arrs[0].shape[2]
    for a in arrs:
        if a.dtype != np.uint8 or (1 if a.ndim == 2 else a.shape[2]) != ch or ch not in (1, 3):
            raise ValueError("images must be uint8 [H, W] or [H, W, 3] with one channel count per batch")
    sizes = [a.size for a in arrs]
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    packed = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs])).pin_memory().to(dev, non_blocking=True)
    b = len(arrs)
    out = torch.empty(b, ch, height, width, dtype=torch.float32, device=dev)
    hs = (ctypes.c_int32 * b)(*[a.shape[0] for a in arrs])
    ws = (ctypes.c_int32 * b)(*[a.shape[1] for a in arrs])
    offs = (ctypes.c_int64 * b)(*offsets.tolist())
    mean, std = (ctypes.c_float * 3)(*MEAN), (ctypes.c_float * 3)(*STD)
    with torch.cuda.device(dev):
        rc = lib.frx_preprocess_u8(ctypes.c_void_p(packed.data_ptr()), offs, hs, ws, b, ch, height, width, mean, std,
                                   1 if rotate_tall else 0, ctypes.c_void_p(out.data_ptr()),
                                   ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("frx_preprocess_u8 failed with code %d" % rc)
    return out
