"""GPU: size-independent properties at BASELINE.json-sized inputs (base 64, 256x256 and 512x512), where the CPU oracle is
too slow to run: run-to-run determinism of a training step, batch-composability of the eval forward (an image's
probabilities do not depend on what else is in the batch -- exercises every tile / chunk / split-K index computation at
full size), probability range, and the integer identity TP+FP+FN+TN = H*W with counts equal to a torch recount."""
import pytest
import torch

from oracle import robust_unet_ref as R

pytestmark = pytest.mark.gpu


def _model(nc, dev, seed=0):
    import rbunet
    torch.manual_seed(seed)
    return rbunet.RobustUNet(nc, 1, 64).to(dev)


@pytest.mark.parametrize("nc,B,S", [(3, 6, 256), (4, 2, 512)])
def test_train_step_is_deterministic(nc, B, S):
    import rbunet
    dev = torch.device("cuda:0")
    x, y = R.synthetic_inputs(B, nc, S, S, seed=11, blobby=True)
    x, y = x.to(dev), y.to(dev)
    masks = R.synthetic_drop_masks(B, 64, seed=3)
    runs = []
    for _ in range(2):
        model = _model(nc, dev).train()
        model.engine.drop_mask_fn = lambda nm, N, C: masks[nm]
        crit = rbunet.RobustBCEDiceLoss(1.0, 0.5 if nc == 4 else 0.0)
        p = model(x)
        loss = crit(p, y)
        loss.backward()
        torch.cuda.synchronize()
        runs.append((p.detach().clone(), loss.item(), [q.grad.clone() for q in model.parameters()], crit.last_counts.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and runs[0][1] == runs[1][1]
    for a, b in zip(runs[0][2], runs[1][2]):
        assert torch.equal(a, b)                                   # fixed-order reductions: bit-identical gradients
    p, counts = runs[0][0], runs[0][3]
    assert torch.isfinite(p).all() and (p >= 0).all() and (p <= 1).all()
    assert (counts.sum(1) == S * S).all()
    pb, yb = p.view(B, -1) > 0.5, y.view(B, -1) > 0.5
    want = torch.stack([(pb & yb).sum(1), (pb & ~yb).sum(1), (~pb & yb).sum(1), (~pb & ~yb).sum(1)], 1)
    assert torch.equal(counts, want)
    for gr in runs[0][2]:
        assert torch.isfinite(gr).all()


def test_eval_forward_is_batch_composable():
    dev = torch.device("cuda:0")
    model = _model(3, dev).eval()
    x, _ = R.synthetic_inputs(5, 3, 256, 256, seed=21, blobby=True)
    x = x.to(dev)
    with torch.no_grad():
        full = model(x)
        for i in (0, 3):
            single = model(x[i:i + 1])
            assert torch.equal(single[0], full[i]), i               # per-image independence in eval mode, bit-exact
        pair = model(x[1:3])
        assert torch.equal(pair, full[1:3])


def test_inference_1024_tile_runs_and_counts():
    import rbunet
    dev = torch.device("cuda:0")
    model = _model(3, dev).eval()
    x, y = R.synthetic_inputs(2, 3, 1024, 1024, seed=5, blobby=True)
    with torch.no_grad():
        p = model(x.to(dev))
    counts = rbunet.confusion_counts(p, y.to(dev))
    assert p.shape == (2, 1, 1024, 1024) and torch.isfinite(p).all()
    assert (counts.sum(1) == 1024 * 1024).all()
    m = rbunet.batch_metrics(p, y.to(dev))
    assert all(0.0 <= d["iou"] <= 1.0 and 0.0 <= d["accuracy"] <= 1.0 for d in m)
