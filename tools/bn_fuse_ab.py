"""Same-process A/B of Engine.fuse_bn_stats (BatchNorm statistics of the 1x1 / stem / dilated convolutions from the generic
kernel's staged epilogue instead of separate passes) on the training step.  Usage: python tools/bn_fuse_ab.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from tools.synthetic import synthetic_batch  # noqa: E402
from tools.tile_stats_ab import timed  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 64).to(dev).train()
    crit = rbunet.RobustBCEDiceLoss()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    xb, yb = synthetic_batch(64, 3, 256, 256, seed=123)
    xb, yb = xb.to(dev), yb.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()

    for rnd in range(3):
        for fuse in (True, False):
            model.engine.fuse_bn_stats = fuse
            for _ in range(3):
                step()
            t = timed(step, reps, flush)
            print(f"step  fuse_bn_stats {'on ' if fuse else 'off'}: {t:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
