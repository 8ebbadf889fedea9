"""Diagnostic (GPU): AttentionGate forward/backward at product shapes against an fp64 CPU evaluation of the reference
formulas (Main_Final.py:143-148) on the same bf16-rounded inputs -- there is no max-routing inside the gate, so the
device should agree to storage-rounding level.  Prints per-gradient relative errors and, for the scalar BatchNorm of
psi, the sums recomputed on the host from the device's own dq / q0 maps (separates the reduction from its inputs)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rbunet import View                                     # noqa: E402
from rbunet.engine import Engine                            # noqa: E402
from rbunet.model import AttentionGate                      # noqa: E402
from gpu_util import bf16r, from_view, rel_l2, to_view      # noqa: E402
from oracle import robust_unet_ref as R                     # noqa: E402


def run(N, C, Fi, H, W, seed, init):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    gate = AttentionGate(C, C, Fi)
    if init == "kaiming":
        for m in gate.modules():
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    gate.to(dev).train()
    gen = torch.Generator().manual_seed(seed)
    g = bf16r(F.relu(torch.randn((N, C, H, W), generator=gen)))
    x = bf16r(F.relu(torch.randn((N, C, H, W), generator=gen) + 0.3))
    da = bf16r(torch.randn((N, C, H, W), generator=gen) * 1e-3)
    eng = Engine(None)
    eng.overlap_wgrad = False
    cat = torch.zeros((N, H, W, 2 * C), dtype=torch.bfloat16, device=dev)
    cat[..., C:] = g.permute(0, 2, 3, 1).to(torch.bfloat16).to(dev)
    st = eng.ag_forward(gate, View(cat, C, C), to_view(x, dev), View(cat, 0, C), N, H, W, True)
    grads = {}
    dskip = View(torch.empty((N, H, W, C), dtype=torch.bfloat16, device=dev))
    dgup = View(torch.zeros((N, H, W, C), dtype=torch.bfloat16, device=dev))
    eng.ag_backward(gate, st, to_view(da, dev), dskip, dgup, grads, "a")
    torch.cuda.synchronize()
    # fp64 reference
    sd = {"a." + k: v.detach().cpu().double() for k, v in gate.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    gq, xq = g.double().requires_grad_(True), x.double().requires_grad_(True)
    out = R.attention_gate(sd, "a", gq, xq, True)
    out.backward(da.double())
    print(f"--- N={N} C={C} F={Fi} {H}x{W} init={init}")
    print(f"out {rel_l2(from_view(View(cat, 0, C)), out.detach()):.3e}  dskip {rel_l2(from_view(dskip), xq.grad):.3e}  "
          f"dg {rel_l2(from_view(dgup), gq.grad):.3e}")
    for k, v in grads.items():
        ref = sd[k].grad
        if k.endswith("0.bias"):
            continue
        print(f"  {k:22s} rel {rel_l2(v.cpu().reshape(ref.shape), ref):.3e}   |ref| {ref.norm().item():.3e}")
    # forward statistics of the scalar BatchNorm
    with torch.no_grad():
        g1 = F.batch_norm(F.conv2d(gq, sd["a.W_g.0.weight"], sd["a.W_g.0.bias"]), None, None, sd["a.W_g.1.weight"], sd["a.W_g.1.bias"], True)
        x1 = F.batch_norm(F.conv2d(xq, sd["a.W_x.0.weight"], sd["a.W_x.0.bias"]), None, None, sd["a.W_x.1.weight"], sd["a.W_x.1.bias"], True)
        q0 = F.conv2d(F.relu(g1 + x1), sd["a.psi.0.weight"], sd["a.psi.0.bias"])
        q0d = st["q0"].cpu().double().reshape(q0.shape[0], H, W)
        print(f"  q0 rel {rel_l2(q0d, q0[:, 0]):.3e}  mean/std ref {q0.mean().item():.4f}/{q0.std().item():.4f}  "
              f"device stats mean {st['stats'][2].item():.4f} rstd {st['stats'][3].item():.4f} (ref rstd {1 / (q0.var(unbiased=False) + 1e-5).sqrt().item():.4f})")
        psi_d = st["psi"].cpu().double().reshape(N, H, W)
        qh = (q0[:, 0] - q0.mean()) / (q0.var(unbiased=False) + 1e-5).sqrt()
        psi_r = torch.sigmoid(sd["a.psi.1.weight"].detach() * qh + sd["a.psi.1.bias"].detach())
        print(f"  psi rel {rel_l2(psi_d, psi_r):.3e}")
        dpsi = (da.double() * x.double()).sum(1)
        dq_r = dpsi * psi_r * (1 - psi_r)
        print(f"  host fp64: dbeta_psi {dq_r.sum().item():.6e} dgamma_psi {(dq_r * qh).sum().item():.6e};  device: "
              f"{grads['a.psi.1.bias'].item():.6e} {grads['a.psi.1.weight'].item():.6e};  autograd ref: "
              f"{sd['a.psi.1.bias'].grad.item():.6e} {sd['a.psi.1.weight'].grad.item():.6e}")
        print(f"  sum|dq| {dq_r.abs().sum().item():.3e}  sum|dq*qhat| {(dq_r * qh).abs().sum().item():.3e}")


if __name__ == "__main__":
    run(4, 64, 32, 256, 256, 1, "kaiming")
    run(4, 128, 64, 128, 128, 2, "kaiming")
    run(2, 64, 32, 64, 64, 3, "default")
    run(2, 512, 256, 8, 8, 4, "kaiming")
