"""GPU parity: rbunet.UNet (the plain 2-class U-Net of train_water_segmentation.py / predict_coastline.py, SURVEY.md §8f
row 2) against the golden vectors produced by the unmodified reference class and the CPU oracle (fp32 and bf16-storage
modes) -- same tolerance scheme as tests/test_gpu_model.py; argmax confusion counts are bit-exact on the device logits."""
import os

import numpy as np
import pytest
import torch

from gpu_util import Report, rel_l2
from oracle import robust_unet_ref as R
from oracle import unet_ref as U

pytestmark = pytest.mark.gpu


def _setup(golden_dir):
    import rbunet
    g = np.load(os.path.join(golden_dir, "unet_c3_32x32.npz"))
    sd = R.synthetic_state_dict(U.unet_shapes(3, 2), seed=11)
    for k in sd:
        if k.endswith(".weight") and sd[k].dim() == 1:
            sd[k] = sd[k].abs()
    x, y = R.synthetic_inputs(2, 3, 32, 32, seed=321, blobby=True)
    dev = torch.device("cuda:0")
    model = rbunet.UNet(3, 2)
    assert list(model.state_dict().keys()) == list(U.unet_shapes(3, 2).keys())
    model.load_state_dict(sd)
    model.to(dev)
    return g, sd, x, y[:, 0].long(), model, dev


def test_unet_eval_forward(golden_dir):
    import rbunet
    g, sd, x, t, model, dev = _setup(golden_dir)
    model.eval()
    with torch.no_grad():
        z = model(x.to(dev))
        zq = U.unet_forward(sd, x, training=False, st=R.BF16)
        sp = {k: (v * (1 + 1e-6 * torch.randn(v.shape, generator=torch.Generator().manual_seed(1)))
                  if v.is_floating_point() else v) for k, v in sd.items()}
        zq2 = U.unet_forward(sp, x, training=False, st=R.BF16)     # the oracle's own sensitivity to round-off
    ref = torch.from_numpy(g["logits_eval"])
    rep = Report()
    rep.check("logits vs fp32 golden", z, ref, 1.5 * rel_l2(zq, ref) + 5e-3)
    rep.check("logits vs bf16-storage oracle", z, zq, 1.5 * rel_l2(zq2, zq) + 5e-3)
    rep.finish()
    crit = rbunet.CrossEntropyArgmaxLoss()
    loss = crit(z, t.to(dev))
    assert abs(loss.item() - U.ce_loss(z.cpu(), t).item()) < 1e-5 * max(1.0, loss.item())       # the loss kernel itself
    assert abs(loss.item() - float(g["loss_eval"])) < 3e-2 * abs(float(g["loss_eval"]))
    counts = crit.last_counts.cpu().numpy()
    assert (counts == U.argmax_counts(z.cpu().numpy(), t.numpy())).all()                         # bit-exact integers
    acc, iou = crit.batch_accuracy_iou()
    assert (acc, iou) == U.batch_accuracy_iou(counts)
    assert abs(acc - float(g["accuracy_eval"])) < 0.02 and abs(iou - float(g["iou_eval"])) < 0.03


def test_unet_train_step(golden_dir):
    import rbunet
    g, sd, x, t, model, dev = _setup(golden_dir)
    model.train()
    crit = rbunet.CrossEntropyArgmaxLoss()
    z = model(x.to(dev))
    loss = crit(z, t.to(dev))
    loss.backward()
    torch.cuda.synchronize()
    names = [n for n, _ in model.named_parameters()]

    def oracle(st, perturb=0.0):
        s = {k: v.clone() for k, v in sd.items()}
        if perturb:
            gen = torch.Generator().manual_seed(1)
            for n in names:
                s[n].mul_(1 + perturb * torch.randn(s[n].shape, generator=gen))
        for n in names:
            s[n].requires_grad_(True)
        nb = {}
        zz = U.unet_forward(s, x, training=True, new_buffers=nb, st=st)
        U.ce_loss(zz, t).backward()
        return zz.detach(), {n: s[n].grad for n in names}, nb

    zf, gf, nbf = oracle(R.FP32)
    zq, gq, _ = oracle(R.BF16)
    zq2, gq2, _ = oracle(R.BF16, 1e-6)
    assert rel_l2(zf, torch.from_numpy(g["logits_train"])) < 1e-4
    rep = Report()
    rep.check("logits vs fp32 oracle", z.detach(), zf, 1.5 * rel_l2(zq, zf) + 5e-3)
    rep.check("logits vs bf16-storage oracle", z.detach(), zq, 1.5 * rel_l2(zq2, zq) + 5e-3)
    assert abs(loss.item() - float(g["loss_train"])) < 3e-2 * abs(float(g["loss_train"]))
    dev_f, q_f, dev_q, q2_q = [], [], [], []
    for n, prm in model.named_parameters():
        got = prm.grad.cpu()
        if n.endswith(".0.bias") or n.endswith(".3.bias"):            # conv bias in front of a train-mode BN
            assert got.abs().max() <= 1e-6 and gf[n].abs().max() < 1e-3, n
            continue
        dev_f.append(rel_l2(got, gf[n])); q_f.append(rel_l2(gq[n], gf[n]))
        dev_q.append(rel_l2(got, gq[n])); q2_q.append(rel_l2(gq2[n], gq[n]))
    rms = lambda v: float(np.sqrt(np.mean(np.square(v))))
    rep.rows.append(("RMS grad deviation vs fp32 oracle", rms(dev_f), 1.5 * rms(q_f) + 0.01))
    rep.rows.append(("RMS grad deviation vs bf16-storage oracle", rms(dev_q), 1.5 * rms(q2_q) + 0.01))
    msd = model.state_dict()
    for k, v in nbf.items():
        if v.dtype == torch.int64:
            assert int(msd[k]) == int(v), k
        else:
            rep.check(k, msd[k], v, 3e-2)
    rep.finish()


def test_unet_works_with_torch_cross_entropy_and_trains():
    import rbunet
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rbunet.UNet(3, 2).to(dev).train()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-3)                    # train_water_segmentation.py:305 uses Adam(1e-4)
    crit = torch.nn.CrossEntropyLoss()                                       # the reference's own criterion on our logits
    x, y = R.synthetic_inputs(4, 3, 64, 64, seed=9, blobby=True)
    x, t = x.to(dev), y[:, 0].long().to(dev)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = crit(model(x), t)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses[-1]) and losses[-1] < 0.7 * losses[0], (losses[0], losses[-1])
    with pytest.raises(RuntimeError):
        rbunet.UNet(3, 2)(torch.zeros((1, 3, 32, 32)))                       # CPU tensors: no fallback
