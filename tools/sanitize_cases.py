"""Small shapes of the four tcgen05 / TMA / mbarrier kernels (conv_halo, conv_gemm, wgrad_halo, wgrad_gemm) plus one tiny
whole-model training step, for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_cases.py

Results are checked against torch CPU convolutions so a run that "passes" the sanitizer also computed the right thing."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rbunet                                   # noqa: E402
from rbunet import ops                          # noqa: E402
from rbunet.engine import Engine                # noqa: E402
from gpu_util import bf16r, from_view, rel_l2, to_view   # noqa: E402


def rnd(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def conv_case(N, H, W, Cin, Cout, ksz, dil, dev):
    x = bf16r(rnd((N, Cin, H, W), 1)).requires_grad_(True)
    w = bf16r(rnd((Cout, Cin, ksz, ksz), 2, (1.0 / (Cin * ksz * ksz)) ** 0.5)).requires_grad_(True)
    dy = bf16r(rnd((N, Cout, H, W), 3))
    y = F.conv2d(x, w, padding=dil * (ksz // 2), dilation=dil)
    y.backward(dy)
    d = dil if ksz == 3 else 0
    xv, dyv, wd = to_view(x.detach(), dev), to_view(dy, dev), w.detach().to(dev).contiguous()
    yo = ops.View(torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=dev))
    ops.conv_gemm(N, H, W, [(xv, ops.pack_weight(wd, 0), ksz * ksz, d, False)], Cout, yo)
    dxo = ops.View(torch.empty((N, H, W, Cin), dtype=torch.bfloat16, device=dev))
    ops.conv_gemm(N, H, W, [(dyv, ops.pack_weight(wd, 1), ksz * ksz, d, False)], Cin, dxo)
    gw = torch.empty((Cout, Cin, ksz, ksz), device=dev)
    eng = Engine(None)
    eng.overlap_wgrad = False
    eng.wgrad(N, H, W, dyv, xv, ksz * ksz, d, False, gw)
    torch.cuda.synchronize()
    e = (rel_l2(from_view(yo), y.detach()), rel_l2(from_view(dxo), x.grad), rel_l2(gw, w.grad))
    assert e[0] < 4e-3 and e[1] < 4e-3 and e[2] < 1e-4, e
    return e


def main():
    dev = torch.device("cuda:0")
    cases = [("conv_halo + wgrad_halo 3x3 64->64 32x32", (2, 32, 32, 64, 64, 3, 1)),
             ("conv_halo + wgrad_halo 3x3 128->128 16x16 (block_n 128)", (1, 16, 16, 128, 128, 3, 1)),
             ("conv_gemm + wgrad_gemm 1x1 64->32", (2, 16, 16, 64, 32, 1, 1)),
             ("conv_gemm + wgrad_gemm 3x3 d2 64->64", (1, 16, 16, 64, 64, 3, 2))]
    for name, c in cases:
        print(name, conv_case(*c, dev), flush=True)
    # one tiny whole-model training step: every kernel of the path at least once
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 16).to(dev).train()
    model.engine.overlap_wgrad = os.environ.get("RBU_SANITIZE_OVERLAP") == "1"
    x = rnd((2, 3, 32, 32), 5).to(dev)
    y = (rnd((2, 1, 32, 32), 6) > 0).float().to(dev)
    crit = rbunet.RobustBCEDiceLoss(1.0, 0.5)
    loss = crit(model(x), y)
    loss.backward()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    print("model step ok, loss", loss.item(), flush=True)


if __name__ == "__main__":
    main()
