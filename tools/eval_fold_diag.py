"""Diagnostic (GPU): eval-mode ResidualBlock forward, folded path (BatchNorm1 + ReLU in conv1's epilogue, no saved state)
and unfolded path, against the bf16-storage oracle at small spatial sizes; and the scale/bias/ReLU conv epilogue alone."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rbunet                                                    # noqa: E402
from rbunet import ops                                           # noqa: E402
from rbunet.engine import Engine                                 # noqa: E402
from rbunet.model import ResidualBlock                           # noqa: E402
from gpu_util import bf16r, from_view, rel_l2, to_view           # noqa: E402
from oracle import robust_unet_ref as R                          # noqa: E402
import test_gpu_widths as T                                      # noqa: E402

dev = torch.device("cuda:0")
for (N, Cin, Cout, H, W) in ((1, 128, 256, 1, 1), (3, 128, 256, 1, 2), (1, 64, 128, 2, 2), (2, 32, 64, 4, 4), (2, 32, 64, 8, 8),
                             (2, 32, 64, 16, 16), (2, 256, 256, 1, 1), (2, 64, 32, 2, 2)):
    sd = R.synthetic_state_dict(T._rb_shapes(Cin, Cout), seed=11)
    x = F.relu(T._rand((N, Cin, H, W), 5))
    q = {"b." + k: v for k, v in sd.items()}
    with torch.no_grad():
        ref = R.residual_block(q, "b", R.BF16.act(bf16r(x)), False, None, st=R.BF16)
    res = []
    for saving in (True, False):
        blk = ResidualBlock(Cin, Cout, 0.1)
        blk.load_state_dict(sd)
        blk.to(dev).eval()
        eng = Engine(None)
        eng._saving = saving
        out, _ = eng.rb_forward("b", blk, to_view(x, dev), N, H, W, False)
        torch.cuda.synchronize()
        res.append(rel_l2(from_view(out), ref))
    # conv + scale + bias + relu alone
    w = bf16r(T._rand((Cout, Cin, 3, 3), 2, (1.0 / (Cin * 9)) ** 0.5))
    sc, bi = T._rand((Cout,), 3).abs() + 0.5, T._rand((Cout,), 4)
    want = F.relu(F.conv2d(bf16r(x), w, padding=1) * sc.view(1, -1, 1, 1) + bi.view(1, -1, 1, 1))
    yo = ops.View(torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device=dev))
    ops.conv_gemm(N, H, W, [(to_view(x, dev), ops.pack_weight(w.to(dev).contiguous(), 0), 9, 1, False)], Cout, yo,
                  scale=sc.to(dev), bias=bi.to(dev), relu=Cout)
    torch.cuda.synchronize()
    print(f"N={N} {Cin}->{Cout} {H}x{W}: block eval unfolded {res[0]:.3e}  folded {res[1]:.3e}   conv+scale+bias+relu {rel_l2(from_view(yo), want):.3e}",
          flush=True)
