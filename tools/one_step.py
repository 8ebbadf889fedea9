"""Profiling driver: a few training steps of the bench workload (no timing, no baselines) for `ncu`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from tools.synthetic import synthetic_batch  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 64).to(dev).train()
    crit = rbunet.RobustBCEDiceLoss()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    x, y = synthetic_batch(B, 3, S, S, seed=123)
    x, y = x.to(dev), y.to(dev)
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print("loss", loss.item(), "launches", rbunet._lib.launch_count())


if __name__ == "__main__":
    main()
