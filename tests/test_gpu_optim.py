"""GPU parity: rbunet.FusedAdam (one multi-tensor launch) against torch.optim.Adam with the reference's settings
(Main_Final.py:552: lr 1e-4, weight_decay 1e-4, coupled L2) over several steps on tensors of ragged sizes."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam():
    import rbunet
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 3), (64,), (1,), (1024, 37), (5, 7, 11), (2049,)]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-4)
    o_mine = rbunet.FusedAdam(mine, lr=1e-2, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(o_mine, mode="min", patience=0, factor=0.5)   # Main_Final.py:553
    for step in range(5):
        for a, b in zip(ref, mine):
            grad = torch.randn(a.shape, generator=g).to(dev)
            a.grad = grad.clone()
            b.grad = grad.clone()
        o_ref.step()
        o_mine.step()
        if step == 2:
            sched.step(1.0)
            sched.step(2.0)                     # no improvement -> lr halves; mirror it on the torch optimizer
            for gr in o_ref.param_groups:
                gr["lr"] = o_mine.param_groups[0]["lr"]
    torch.cuda.synchronize()
    assert o_mine.param_groups[0]["lr"] == pytest.approx(5e-3)
    for a, b in zip(ref, mine):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (a - b).abs().max().item()
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_fused_adam_trains_the_model():
    import numpy as np
    import rbunet
    from oracle import robust_unet_ref as R
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 16).to(dev).train()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    crit = rbunet.RobustBCEDiceLoss()
    x, y = R.synthetic_inputs(4, 3, 64, 64, seed=9, blobby=True)
    x, y = x.to(dev), y.to(dev)
    losses = []
    for _ in range(40):
        opt.zero_grad()
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses[-1]) and losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
