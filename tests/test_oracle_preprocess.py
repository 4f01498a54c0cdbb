"""The input-pipeline oracle (oracle/preprocess.py) pinned against cv2.resize itself -- the library the reference's
A.Resize calls (data/augmentations.py:27-46)."""
import numpy as np
import pytest

from oracle import preprocess

cv2 = pytest.importorskip("cv2")


def test_resize_bit_exact_with_cv2():
    rng = np.random.default_rng(0)
    shapes = [(256, 512), (256, 512, 3), (128, 256), (37, 911, 3), (400, 90), (8, 8, 3), (129, 257), (1, 40), (300, 1, 3)]
    shapes += [(int(rng.integers(2, 400)), int(rng.integers(2, 900))) + ((3,) if i % 2 else ()) for i in range(120)]
    for shp in shapes:
        img = rng.integers(0, 256, shp, dtype=np.uint8)
        want = cv2.resize(img, (256, 128), interpolation=cv2.INTER_LINEAR)
        got = preprocess.resize_linear_u8(img, 128, 256)
        assert np.array_equal(got, want), shp


def test_load_item_layout_and_rotation():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (300, 100, 3), dtype=np.uint8)       # h / w > 2: rotated by 90 degrees first
    out = preprocess.load_item(img, 128, 256)
    assert out.shape == (3, 128, 256) and out.dtype == np.float32
    rot = cv2.rotate(img, cv2.ROTATE_90_COUNTERCLOCKWISE)
    want = cv2.resize(rot, (256, 128), interpolation=cv2.INTER_LINEAR).astype(np.float32)
    want = (want - np.array(preprocess.MEAN, np.float32) * 255) * (1 / (np.array(preprocess.STD, np.float32) * 255)).astype(np.float32)
    assert np.abs(out - want.transpose(2, 0, 1)).max() <= 1e-6
