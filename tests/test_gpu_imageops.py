"""GPU parity (bit-exact: integer / byte work) of rbu_enhance_image and rbu_coastline_mask through the C ABI against the
numpy oracle and the committed reference / OpenCV fixtures."""
import os

import numpy as np
import pytest
import torch

import rbunet
from oracle import imageops_ref as I

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "imageops.npz"))
DEV = "cuda:0"


def _dev(a):
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).to(DEV).view(torch.uint16)
    return torch.from_numpy(a).to(DEV)


@pytest.mark.parametrize("tag", ["u16", "u8", "u16full", "tiny"])
def test_enhance_fixture(tag):
    rgb = GOLD[f"enh_{tag}_in"]
    for ew in (0, 1):
        out, pct = rbunet.enhance_image(_dev(rgb), bool(ew), return_percentiles=True)
        assert np.array_equal(out.cpu().numpy(), GOLD[f"enh_{tag}_out{ew}"])
        assert np.array_equal(pct.cpu().numpy(), GOLD[f"enh_{tag}_pct"])      # float64 percentiles, bit for bit


@pytest.mark.parametrize("dt,hi,shape", [(np.uint16, 30000, (2, 300, 257, 3)), (np.uint8, 256, (3, 129, 64, 4)),
                                         (np.uint16, 65536, (1, 64, 64, 6)), (np.uint8, 256, (1, 1, 1, 1))])
def test_enhance_random_batches(dt, hi, shape):
    rng = np.random.default_rng(11)
    x = np.clip(rng.gamma(2.0, hi / 8.0, size=shape), 0, hi - 1).astype(dt)
    got = rbunet.enhance_image(_dev(x), True).cpu().numpy()
    for b in range(shape[0]):
        assert np.array_equal(got[b], I.enhance_image(x[b], True)), b


def test_enhance_constant_band():
    x = np.full((16, 16, 3), 500, dtype=np.uint16)
    x[:, :, 1] = np.arange(256, dtype=np.uint16).reshape(16, 16) * 7
    assert np.array_equal(rbunet.enhance_image(_dev(x), True).cpu().numpy(), I.enhance_image(x, True))


def test_enhance_full_size_properties():
    """1024^2 x 4 bands uint16: agrees with numpy's percentile and is monotone in the input (size-independent checks)."""
    rng = np.random.default_rng(1)
    x = np.clip(rng.gamma(2.0, 3000.0, size=(1024, 1024, 4)), 0, 65535).astype(np.uint16)
    out, pct = rbunet.enhance_image(_dev(x), False, return_percentiles=True)
    out, pct = out.cpu().numpy(), pct.cpu().numpy()
    for c in range(4):
        assert pct[c, 0] == np.percentile(x[:, :, c], 2) and pct[c, 1] == np.percentile(x[:, :, c], 98)
        order = np.argsort(x[:, :, c].reshape(-1), kind="stable")
        assert (np.diff(out[:, :, c].reshape(-1)[order].astype(np.int16)) >= 0).all()
        assert out[:, :, c].min() == 0 and out[:, :, c].max() == 255


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_coastline_fixture(tag):
    mask = GOLD[f"mask_{tag}"]
    for k in (5, 20, 3):
        assert np.array_equal(rbunet.coastline_mask(_dev(mask), k).cpu().numpy(), GOLD[f"coast_{tag}_k{k}"])
        assert np.array_equal(rbunet.coastline_mask(_dev(mask // 255), k).cpu().numpy(), GOLD[f"coast01_{tag}_k{k}"])


@pytest.mark.parametrize("shape", [(3, 70, 45), (2, 37, 64), (1, 9, 260)])   # byte-wise path / packed-word path (W % 4 == 0)
@pytest.mark.parametrize("k", [1, 2, 5, 20, 33, 64])
def test_coastline_random_batch(k, shape):
    rng = np.random.default_rng(k)
    m = rng.integers(0, 256, size=shape).astype(np.uint8)          # general uint8 values, ragged tile edges
    m[0] = (m[0] > 200) * 255
    got = rbunet.coastline_mask(_dev(m), k).cpu().numpy()
    for b in range(shape[0]):
        assert np.array_equal(got[b], I.coastline_mask(m[b], k)), (k, b)


def test_coastline_full_size_properties():
    """1024^2 mask: the band never overlaps the water, and dilating with k=1 gives an empty band."""
    rng = np.random.default_rng(4)
    low = rng.random((130, 130))
    m = ((np.kron(low, np.ones((8, 8)))[:1024, :1024] > 0.5) * 255).astype(np.uint8)
    band = rbunet.coastline_mask(_dev(m), 5).cpu().numpy()
    assert not (band.astype(bool) & m.astype(bool)).any()
    assert set(np.unique(band)) <= {0, 255}
    assert not rbunet.coastline_mask(_dev(m), 1).any().item()
    rows = slice(100, 140)
    assert np.array_equal(band[rows], I.coastline_mask(m, 5)[rows])


def test_errors():
    with pytest.raises(RuntimeError):
        rbunet.enhance_image(torch.zeros((4, 4, 3), dtype=torch.uint8))            # CPU tensor
    with pytest.raises(RuntimeError):
        rbunet.coastline_mask(torch.zeros((4, 4), dtype=torch.uint8, device=DEV), 65)
    with pytest.raises(RuntimeError):
        rbunet.enhance_image(torch.zeros((4, 4, 9), dtype=torch.uint8, device=DEV))
