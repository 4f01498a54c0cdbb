"""Device input pipeline (frx.data.transform_batch) against the oracle restatement of the reference's transform chain
(oracle/preprocess.py, pinned to cv2.resize): bit-exact tensors for ragged batches, both channel counts, tall images
(rotated), exact 2x down-scaling (OpenCV's box-average path) and up-scaling."""
import numpy as np
import pytest
import torch

import frx
from oracle import preprocess

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("channels", [1, 3])
def test_transform_batch_is_bit_exact(channels):
    rng = np.random.default_rng(channels)
    shapes = [(256, 512), (128, 256), (37, 911), (300, 100), (64, 64), (500, 90), (129, 257), (2, 40), (90, 1000), (257, 511)]
    imgs = [rng.integers(0, 256, s + ((3,) if channels == 3 else ()), dtype=np.uint8) for s in shapes]
    out = frx.data.transform_batch(imgs, 128, 256).cpu().numpy()
    assert out.shape == (len(imgs), channels, 128, 256)
    for i, im in enumerate(imgs):
        want = preprocess.load_item(im, 128, 256)
        assert np.array_equal(out[i], want), (i, im.shape, np.abs(out[i] - want).max())


def test_transform_feeds_the_model(ckpt0):
    """End to end: uint8 pages -> device transform -> greedy decode equals decoding the oracle-transformed tensor."""
    from helpers import make_model
    rng = np.random.default_rng(7)
    imgs = [rng.integers(0, 256, (int(rng.integers(60, 300)), int(rng.integers(200, 700))), dtype=np.uint8) for _ in range(3)]
    model = make_model(ckpt0).cuda().eval()
    x_dev = frx.data.transform_batch(imgs, 128, 256)
    x_ref = torch.from_numpy(np.stack([preprocess.load_item(im, 128, 256) for im in imgs])).cuda()
    assert torch.equal(x_dev, x_ref)
    with torch.no_grad():
        _, t1 = model.greedy(x_dev, 12)
        _, t2 = model.greedy(x_ref, 12)
    assert torch.equal(t1, t2)
