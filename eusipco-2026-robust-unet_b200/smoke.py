"""smoke(): one tiny invocation of the hot path on the GPU, checked against the oracle."""
import torch
import torch.nn.functional as F


def run(dev):
    from . import _lib, ops
    _lib.check(_lib.lib().rbu_device_check(), "rbu_device_check")
    from oracle import robust_unet_ref as R
    # conv 3x3 64->64 on 2x16x16 + loss/metrics
    g = torch.Generator().manual_seed(0)
    x = torch.randn((2, 64, 16, 16), generator=g)
    w = torch.randn((64, 64, 3, 3), generator=g) * 0.04
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1)
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev)
    y = torch.empty((2, 16, 16, 64), dtype=torch.bfloat16, device=dev)
    ops.conv_gemm(2, 16, 16, [(ops.View(xb), ops.pack_weight(w.to(dev), 0), 9, 1, False)], 64, ops.View(y))
    out = y.float().cpu().permute(0, 3, 1, 2)
    err = ((out - ref).norm() / ref.norm()).item()
    assert err < 4e-3, f"conv_gemm smoke mismatch: rel-L2 {err}"
    p = torch.sigmoid(torch.randn((2, 1, 16, 16), generator=g) * 3)
    t = (torch.rand((2, 1, 16, 16), generator=g) > 0.5).float()
    loss, _, counts = ops.loss_forward(p.to(dev), t.to(dev))
    assert abs(loss.item() - R.bce_loss(p, t).item()) < 1e-5
    assert (counts.cpu().numpy() == R.confusion_counts(p.numpy(), t.numpy())).all()
    print(f"smoke ok: conv rel-L2 {err:.2e}, loss {loss.item():.6f}")
