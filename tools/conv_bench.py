"""Micro-benchmark of the implicit-GEMM convolution kernels on the Robust U-Net layer shapes (batch 64 at 256x256).
Usage: python tools/conv_bench.py [reps] [which]     which: fwd | wgrad | all"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from rbunet import ops  # noqa: E402
from rbunet.engine import Engine  # noqa: E402

SHAPES = [  # name, H=W, Cin, Cout, taps
    ("L0 3x3 64->64", 256, 64, 64, 9),
    ("L0 3x3 128->64", 256, 128, 64, 9),
    ("L1 3x3 128->128", 128, 128, 128, 9),
    ("L1 3x3 256->128", 128, 256, 128, 9),
    ("L2 3x3 256->256", 64, 256, 256, 9),
    ("L3 3x3 512->512", 32, 512, 512, 9),
    ("L4 3x3 1024->1024", 16, 1024, 1024, 9),
    ("D4 3x3 1024->512", 32, 1024, 512, 9),
    ("D3 3x3 512->256", 64, 512, 256, 9),
    ("L0 1x1 128->64", 256, 128, 64, 1),
    ("L1 1x1 256->128", 128, 256, 128, 1),
    ("L3 1x1 1024->512", 32, 1024, 512, 1),
    ("A4 1x1 512->256", 32, 512, 256, 1),
    ("D3 1x1 512->256", 64, 512, 256, 1),
    ("A3 1x1 256->128", 64, 256, 128, 1),
    ("dil 1x1 512->256", 16, 512, 256, 1),
]


AUX = [  # name, H=W, Cin, Ncols, mode (plain | addend | scatter)  -- the small-K, epilogue-bound GEMMs of the step
    ("stem 1x1 32->128", 256, 32, 128, "plain"),
    ("AG fwd 1x1 64->32", 256, 64, 32, "plain"),
    ("AG dgrad 1x1 32->64 +add", 256, 32, 64, "addend"),
    ("AG dgrad 1x1 64->128 +add", 128, 64, 128, "addend"),
    ("proj 1x1 128->64", 256, 128, 64, "plain"),
    ("up 128->4x64 scatter", 128, 128, 256, "scatter"),
    ("up 256->4x128 scatter", 64, 256, 512, "scatter"),
]


def aux(reps, B, dev, flush):
    for name, S, ci, nc, mode in AUX:
        x = ops.View(torch.randn((B, S, S, ci), device=dev).to(torch.bfloat16))
        w = torch.randn((nc, ci, 1, 1), device=dev) * 0.05
        wp = ops.pack_weight(w, 0)
        kw = {}
        if mode == "scatter":
            co = nc // 4
            y = ops.View(torch.empty((B, 2 * S, 2 * S, co), dtype=torch.bfloat16, device=dev))
            kw = dict(scatter=True, Cout=co)
        else:
            y = ops.View(torch.zeros((B, S, S, nc), dtype=torch.bfloat16, device=dev))
            if mode == "addend":
                kw = dict(addend=y)
        nbytes = 2.0 * B * S * S * (ci + nc * (2 if mode == "addend" else 1))
        ms = []
        for i in range(reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv_gemm(B, S, S, [(x, wp, 1, 0, False)], nc, y, **kw)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(e0.elapsed_time(e1))
        t = sorted(ms)[len(ms) // 2]
        print(f"aux   {name:28s} {t:8.3f} ms  {nbytes / t / 1e6:8.1f} GB/s", flush=True)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    which = sys.argv[2] if len(sys.argv) > 2 else "all"
    B = 64
    dev = torch.device("cuda:0")
    eng = Engine(None)
    eng.overlap_wgrad = False          # time the weight-gradient kernels on the stream the events are recorded on
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if which == "aux":
        return aux(reps, B, dev, flush)
    for name, S, ci, co, taps in SHAPES:
        x = ops.View(torch.randn((B, S, S, ci), device=dev).to(torch.bfloat16))
        y = ops.View(torch.empty((B, S, S, co), dtype=torch.bfloat16, device=dev))
        k = 3 if taps == 9 else 1
        w = torch.randn((co, ci, k, k), device=dev) * 0.05
        wp = ops.pack_weight(w, 0)
        flops = 2.0 * B * S * S * ci * co * taps
        if which in ("fwd", "all"):
            ms = []
            for i in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv_gemm(B, S, S, [(x, wp, taps, 1, False)], co, y)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ms.append(e0.elapsed_time(e1))
            t = sorted(ms)[len(ms) // 2]
            print(f"fwd   {name:22s} {t:8.3f} ms  {flops / t / 1e9:8.1f} TFLOP/s", flush=True)
        if which in ("wgrad", "all"):
            dy = ops.View(torch.randn((B, S, S, co), device=dev).to(torch.bfloat16))
            g = torch.empty_like(w)
            ms = []
            for i in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.wgrad(B, S, S, dy, x, taps, 1, False, g)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    ms.append(e0.elapsed_time(e1))
            t = sorted(ms)[len(ms) // 2]
            print(f"wgrad {name:22s} {t:8.3f} ms  {flops / t / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
