"""Same-process A/B of Engine.tile_stats_eval (ChannelAttention's pooled statistics from the conv2 epilogue in eval mode) on
the 1024 x 1024 inference workload.  Usage: python tools/infer_ab.py [reps] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from tools.synthetic import synthetic_batch  # noqa: E402
from tools.tile_stats_ab import timed  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 64).to(dev).eval()
    xb, _ = synthetic_batch(B, 3, 1024, 1024, seed=123)
    xb = xb.to(dev)

    def fwd():
        with torch.no_grad():
            model(xb)

    for rnd in range(3):
        for on in (True, False):
            model.engine.tile_stats_eval = on
            for _ in range(2):
                fwd()
            t = timed(fwd, reps, flush)
            print(f"inference {B} x 1024^2  tile_stats_eval {'on ' if on else 'off'}: {t:8.3f} ms  {B / t * 1e3:7.1f} img/s", flush=True)


if __name__ == "__main__":
    main()
