"""Diagnostic (GPU): rbu_bn_stats at the shapes of the training step (dense and channel-slice views), for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from rbunet import ops  # noqa: E402
from rbunet.engine import Engine  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    eng = Engine(None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = 64
    for name, S, C, ld, off in (("AG yg dense C=32 @256", 256, 32, 32, 0), ("stem slice C=64 of 128 @256", 256, 64, 128, 0),
                                ("dense C=64 @256", 256, 64, 64, 0), ("dense C=128 @128", 128, 128, 128, 0)):
        buf = torch.randn((B, S, S, ld), device=dev).to(torch.bfloat16)
        x = ops.View(buf, off, C)
        bn = torch.nn.BatchNorm2d(C).to(dev)
        ms = []
        for i in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.bn_stats(x, B, S * S, bn, True)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[2]
        print(f"{name:32s} {t:7.3f} ms  {B * S * S * C * 2 / t / 1e6:7.0f} GB/s (useful bytes)", flush=True)


if __name__ == "__main__":
    main()
