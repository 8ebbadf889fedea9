"""Throughput of the image operations of SURVEY.md §8(f) rows 3-4 (HBM-bound byte kernels) next to the numpy / OpenCV
oracle on one image.  Usage: python tools/imageops_bench.py [batch]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from oracle import imageops_ref as I  # noqa: E402  (CPU baseline leg only)


def timed(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = 1024
    rng = np.random.default_rng(0)
    x = np.clip(rng.gamma(2.0, 3000.0, size=(B, S, S, 3)), 0, 65535).astype(np.uint16)
    xd = torch.from_numpy(x.view(np.int16)).cuda().view(torch.uint16)
    t = timed(lambda: rbunet.enhance_image(xd, True))
    nbytes = x.nbytes * 2 + x.size            # histogram read + stretch read + uint8 write
    print(f"enhance_image  {B}x{S}x{S}x3 uint16: {t:.3f} ms  {B / t * 1e3:.0f} img/s  {nbytes / t / 1e6:.0f} GB/s (algorithmic)")
    t0 = time.perf_counter()
    I.enhance_image(x[0], True)
    t1 = time.perf_counter()
    print(f"  numpy oracle, 1 image: {(t1 - t0) * 1e3:.1f} ms  ({1 / (t1 - t0):.1f} img/s)")
    low = rng.random((B, S // 8 + 1, S // 8 + 1))
    m = ((np.kron(low, np.ones((1, 8, 8)))[:, :S, :S] > 0.5) * 255).astype(np.uint8)
    md = torch.from_numpy(m).cuda()
    for k in (5, 20):
        t = timed(lambda: rbunet.coastline_mask(md, k))
        print(f"coastline_mask {B}x{S}x{S} k={k}: {t:.3f} ms  {B / t * 1e3:.0f} img/s  {2 * m.nbytes / t / 1e6:.0f} GB/s (algorithmic)")
        try:
            import cv2
            kern = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
            t0 = time.perf_counter()
            for _ in range(5):
                _ = cv2.dilate(m[0], kern, iterations=1) - m[0]
            t1 = time.perf_counter()
            print(f"  OpenCV on the host, 1 image: {(t1 - t0) / 5 * 1e3:.2f} ms")
        except ImportError:
            pass


if __name__ == "__main__":
    main()
