"""Debug aid (not a test): run the whole model on the GPU and print per-tensor deviations from the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rbunet  # noqa: E402
from gpu_util import rel_l2  # noqa: E402
from oracle import robust_unet_ref as R  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "model_c3_b16_32x32.npz"
    g = np.load(os.path.join(ROOT, "tests", "golden", name))
    n_channels, base, batch, h, w = [int(v) for v in g["config"]]
    sd = R.synthetic_state_dict(R.robust_unet_shapes(n_channels, 1, base), seed=0)
    x, y = R.synthetic_inputs(batch, n_channels, h, w, seed=123, blobby=True)
    masks = R.synthetic_drop_masks(batch, base, seed=7)
    dev = torch.device("cuda:0")
    model = rbunet.RobustUNet(n_channels, 1, base)
    model.load_state_dict(sd)
    model.to(dev)
    model.eval()
    with torch.no_grad():
        p = model(x.to(dev))
    print("eval probs rel-L2", rel_l2(p, torch.from_numpy(g["probs_eval"])))
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        za = R.robust_unet_forward(sd, x, training=False, return_logits=True)
    print("eval probs rel-L2 of torch bf16 autocast (yardstick)", rel_l2(torch.sigmoid(za.float()), torch.from_numpy(g["probs_eval"])))
    with torch.no_grad():
        zo = R.robust_unet_forward(sd, x, training=False, return_logits=True)
    zd = torch.log(p.cpu().double() / (1 - p.cpu().double()))
    ok = torch.isfinite(zd)
    print("eval logits rel-L2 ours", rel_l2(zd[ok].float(), zo[ok]), "autocast", rel_l2(za.float(), zo), "logit range", zo.min().item(), zo.max().item())
    model.train()
    model.engine.drop_mask_fn = lambda nm, N, C: masks[nm]
    w_dice = float(g["w_dice"])
    crit = rbunet.RobustBCEDiceLoss(1.0, w_dice)
    p = model(x.to(dev))
    loss = crit(p, y.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    print("train probs rel-L2", rel_l2(p.detach(), torch.from_numpy(g["probs_train"])), "loss", loss.item(),
          "ref", float(g["loss_train"]))
    osd = {k: v.clone() for k, v in sd.items()}
    for n, _ in model.named_parameters():
        osd[n].requires_grad_(True)
    po = R.robust_unet_forward(osd, x, training=True, drop_masks=masks)
    R.bce_dice_loss(po, y, 1.0, w_dice).backward()
    # yardstick: torch's own bf16 autocast on the CPU oracle
    asd = {k: v.clone() for k, v in sd.items()}
    for n, _ in model.named_parameters():
        asd[n].requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        pa = R.robust_unet_forward(asd, x, training=True, drop_masks=masks, return_logits=True)
    R.bce_dice_loss(torch.sigmoid(pa.float()), y, 1.0, w_dice).backward()
    qsd = {k: v.clone() for k, v in sd.items()}
    for n, _ in model.named_parameters():
        qsd[n].requires_grad_(True)
    pq = R.robust_unet_forward(qsd, x, training=True, drop_masks=masks, st=R.BF16)
    R.bce_dice_loss(pq, y, 1.0, w_dice).backward()
    print("train probs vs bf16-storage oracle", rel_l2(p.detach(), pq.detach()))
    for n, prm in model.named_parameters():
        ref = osd[n].grad
        got = prm.grad.cpu()
        print(f"{n:40s} ours {rel_l2(got, ref):9.3e}  autocast {rel_l2(asd[n].grad, ref):9.3e}  vs-q {rel_l2(got, qsd[n].grad):9.3e} |ref| {ref.norm().item():.3e}")


if __name__ == "__main__":
    main()
