"""Diagnostic (GPU): gradient cosine / probability deviation of small base-64 configurations against the fp32 CPU oracle
(choosing the configuration and bars of __graft_entry__.smoke())."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet                                   # noqa: E402
from oracle import robust_unet_ref as R        # noqa: E402


def run(init, S, B):
    dev = torch.device("cuda:0")
    if init == "reference":
        torch.manual_seed(0)
        sd = {k: v.clone() for k, v in rbunet.RobustUNet(3, 1, 64).state_dict().items()}
    else:
        sd = R.synthetic_state_dict(R.robust_unet_shapes(3, 1, 64), seed=0)
    x, y = R.synthetic_inputs(B, 3, S, S, seed=123, blobby=True)
    masks = R.synthetic_drop_masks(B, 64, seed=7)
    model = rbunet.RobustUNet(3, 1, 64)
    model.load_state_dict(sd)
    model.to(dev).train()
    model.engine.drop_mask_fn = lambda name, N, C: masks[name]
    crit = rbunet.RobustBCEDiceLoss()
    p = model(x.to(dev))
    loss = crit(p, y.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    t0 = time.time()
    osd = {k: v.clone() for k, v in sd.items()}
    for n, _ in model.named_parameters():
        osd[n].requires_grad_(True)
    po = R.robust_unet_forward(osd, x, training=True, drop_masks=masks)
    lo = R.bce_loss(po, y)
    lo.backward()
    t1 = time.time() - t0
    a = torch.cat([prm.grad.flatten().cpu().double() for _, prm in model.named_parameters()])
    b = torch.cat([osd[n].grad.flatten().double() for n, _ in model.named_parameters()])
    cos = (a @ b / (a.norm() * b.norm())).item()
    ep = ((p.detach().cpu().double() - po.detach().double()).norm() / po.detach().double().norm()).item()
    print(f"init={init:9s} S={S:3d} B={B}: cosine {cos:.4f}  probs rel-L2 {ep:.3e}  loss {loss.item():.5f} vs {lo.item():.5f}  "
          f"(oracle {t1:.1f} s)", flush=True)


if __name__ == "__main__":
    for cfg in (("reference", 64, 2), ("reference", 64, 4), ("reference", 128, 2), ("reference", 128, 4),
                ("synthetic", 64, 2), ("synthetic", 128, 2), ("synthetic", 128, 4)):
        run(*cfg)
