"""Same-process A/B: BatchNorm / pooling statistics from the 3x3 convolution epilogue (tile statistics) against the separate
bn_stats passes, (a) per layer through ops.conv_gemm and (b) on the whole training step.
Usage: python tools/tile_stats_ab.py [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from rbunet import _lib, ops  # noqa: E402
from tools.synthetic import synthetic_batch  # noqa: E402


def timed(fn, reps, flush):
    ms = []
    for i in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
    return sorted(ms)[len(ms) // 2]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = 64
    for name, S, ci, co in [("64->64 @256", 256, 64, 64), ("128->64 @256", 256, 128, 64), ("64->128 @128", 128, 64, 128), ("128->128 @128", 128, 128, 128), ("256->128 @128", 128, 256, 128),
                            ("256->256 @64", 64, 256, 256), ("512->512 @32", 32, 512, 512), ("1024->1024 @16", 16, 1024, 1024)]:
        x = ops.View(torch.randn((B, S, S, ci), device=dev).to(torch.bfloat16))
        y = ops.View(torch.empty((B, S, S, co), dtype=torch.bfloat16, device=dev))
        wp = ops.pack_weight(torch.randn((co, ci, 3, 3), device=dev) * 0.05, 0)
        ts = torch.empty(_lib.lib().rbu_conv_tile_stats_floats(B, S, S, co), dtype=torch.float32, device=dev)
        t0 = timed(lambda: ops.conv_gemm(B, S, S, [(x, wp, 9, 1, False)], co, y), reps, flush)
        t1 = timed(lambda: ops.conv_gemm(B, S, S, [(x, wp, 9, 1, False)], co, y, tile_stats=ts), reps, flush)
        print(f"conv3x3 {name:16s} plain {t0:7.3f} ms   with tile statistics {t1:7.3f} ms   (+{t1 - t0:6.3f})", flush=True)

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
        for fuse, min_c in ((True, 128), (True, 64), (False, 128)):
            model.engine.fuse_tile_stats = fuse
            model.engine.tile_stats_min_c = min_c
            for _ in range(3):
                step()
            t = timed(step, reps, flush)
            print(f"step  tile statistics {'on ' if fuse else 'off'} (from {min_c} channels): {t:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
