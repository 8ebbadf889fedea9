"""GPU parity: fused loss + confusion-count kernels against the oracle (bit-exact integer counts)."""
import numpy as np
import pytest
import torch

from oracle import robust_unet_ref as R

pytestmark = pytest.mark.gpu


def _case(B, H, W, seed, saturate=True):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn((B, 1, H, W), generator=g) * 6
    p = torch.sigmoid(z)
    if saturate:
        p.view(-1)[:7] = torch.tensor([0.0, 1.0, 0.5, 0.5, 1.0, 0.0, 0.49999997])
    y = (torch.rand((B, 1, H, W), generator=g) > 0.5).float()
    return p, y


@pytest.mark.parametrize("B,H,W", [(1, 16, 16), (3, 32, 48), (8, 256, 256), (2, 17, 5)])
@pytest.mark.parametrize("w_dice", [0.0, 0.5])
def test_loss_forward_backward(B, H, W, w_dice):
    from rbunet import ops
    p, y = _case(B, H, W, 1)
    pr = p.clone().requires_grad_(True)
    ref = R.bce_dice_loss(pr, y, 1.0, w_dice, 1.0)
    ref.backward()
    dev = torch.device("cuda:0")
    pd, yd = p.to(dev), y.to(dev)
    loss, sums, counts = ops.loss_forward(pd, yd, 0.5, 1.0, w_dice, 1.0)
    dp = ops.loss_backward(pd, yd, None, sums, 1.0, w_dice, 1.0)
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    assert (counts.cpu().numpy() == R.confusion_counts(p.numpy(), y.numpy())).all()     # bit-exact
    g_ref = pr.grad
    finite = torch.isfinite(g_ref)
    np.testing.assert_allclose(dp.cpu()[finite].numpy(), g_ref[finite].numpy(), rtol=2e-5, atol=1e-9)


def test_counts_edge_cases(golden_dir):
    import os
    from rbunet import ops
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    dev = torch.device("cuda:0")
    counts = ops.confusion_counts(torch.from_numpy(g["pred"]).to(dev), torch.from_numpy(g["target"]).to(dev))
    torch.cuda.synchronize()
    c = counts.cpu().numpy()
    assert (c == R.confusion_counts(g["pred"], g["target"])).all()
    for i in range(4):
        m = R.metrics_from_counts(*c[i])
        for j, k in enumerate(g["metric_keys"]):
            assert abs(m[str(k)] - g["metrics"][i, j]) < 1e-12


def test_empty_input_is_an_error():
    from rbunet import ops
    dev = torch.device("cuda:0")
    e = torch.zeros((0, 1, 4, 4), device=dev)
    with pytest.raises((RuntimeError, ZeroDivisionError)):
        ops.confusion_counts(e, e)
