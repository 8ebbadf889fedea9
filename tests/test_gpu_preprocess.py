"""GPU parity: rbu_preprocess (Normalize of Main_Final.py:697-701 + build-defined HSV planes) is BIT-IDENTICAL to the
numpy oracle on uint8 inputs, including grey / black / saturated pixels."""
import numpy as np
import pytest
import torch

from oracle import robust_unet_ref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nc", [3, 4, 6])
def test_preprocess_bit_exact(nc):
    import rbunet
    rng = np.random.RandomState(nc)
    img = rng.randint(0, 256, size=(3, 37, 29, 3)).astype(np.uint8)
    img[0, 0, :6] = [[0, 0, 0], [255, 255, 255], [128, 128, 128], [255, 0, 0], [0, 255, 0], [0, 0, 255]]
    img[0, 1, :3] = [[10, 200, 200], [200, 10, 200], [200, 200, 10]]
    want = R.preprocess(img, nc)
    got = rbunet.preprocess(torch.from_numpy(img).cuda(), nc).cpu().numpy()
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_preprocess_rejects_bad_input():
    import rbunet
    with pytest.raises(RuntimeError):
        rbunet.preprocess(torch.zeros((1, 4, 4, 3), dtype=torch.uint8), 3)           # CPU tensor
    with pytest.raises(RuntimeError):
        rbunet.preprocess(torch.zeros((1, 4, 4, 3), dtype=torch.uint8).cuda(), 5)    # unsupported channel count
