"""Micro-benchmark of the bandwidth-bound kernels on Robust U-Net tensor shapes (batch 64 at 256x256): algorithmic
GB/s (distinct tensor bytes read + written, each once) against the measured HBM copy peak."""
import json
import os
import sys
from ctypes import c_void_p

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet  # noqa: E402
from rbunet import _lib, ops  # noqa: E402
from rbunet._lib import call, stream_ptr  # noqa: E402
from rbunet.engine import Engine  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps, flush):
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
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    dev = torch.device("cuda:0")
    B = 64
    eng = Engine(None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    p = lambda t: c_void_p(t.data_ptr())
    for S, C in ((256, 64), (128, 128), (32, 512)):
        P, HW = B * S * S, S * S
        tb = P * C * 2 / 1e9          # GB per bf16 tensor
        x = ops.View(torch.randn((B, S, S, C), device=dev).to(torch.bfloat16))
        y = ops.View(torch.empty((B, S, S, C), dtype=torch.bfloat16, device=dev))
        z = ops.View(torch.randn((B, S, S, C), device=dev).to(torch.bfloat16))
        bn = torch.nn.BatchNorm2d(C).to(dev)
        rows = []
        st = {}

        def stats(pool):
            st.update(eng.bn_stats(x, B, HW, bn, True, pool=pool))
        rows.append(("bn_stats", timeit(lambda: stats(False), reps, flush), 1 * tb))
        rows.append(("bn_stats+pool", timeit(lambda: stats(True), reps, flush), 1 * tb))
        drop = torch.ones((B, C), device=dev)
        rows.append(("affine_act", timeit(lambda: call("rbu_affine_act", c_void_p(x.ptr), x.ld, c_void_p(y.ptr), y.ld, P, HW,
                                                        C, p(st["scale"]), p(st["shift"]), p(drop), 1, stream_ptr()), reps, flush), 2 * tb))
        a2g = torch.rand((B, C), device=dev)
        gs = torch.rand(P, device=dev)
        rows.append(("rb_out(proj)", timeit(lambda: call("rbu_rb_out", c_void_p(x.ptr), x.ld, c_void_p(z.ptr), z.ld, c_void_p(y.ptr),
                                                          y.ld, P, HW, C, p(a2g), p(a2g), p(gs), p(st["scale"]), p(st["shift"]),
                                                          stream_ptr()), reps, flush), 3 * tb))
        sums = torch.empty(2 * C, device=dev)
        ws = eng.bwd_ws(B, HW, C, dev)
        rows.append(("bn_bwd(2 passes)", timeit(lambda: call("rbu_bn_bwd", c_void_p(z.ptr), z.ld, c_void_p(x.ptr), x.ld, c_void_p(y.ptr),
                                                              y.ld, B, HW, C, p(st["scale"]), p(st["shift"]), p(st["mean"]), p(st["rstd"]),
                                                              p(drop), 1, p(sums), p(ws), ws.numel() * 4, stream_ptr()), reps, flush), 5 * tb))
        amax = torch.zeros(P, dtype=torch.int32, device=dev)
        ncarg = torch.zeros((B, C), dtype=torch.int32, device=dev)
        s_out = torch.empty((P, 2), device=dev)
        rows.append(("sa_reduce", timeit(lambda: call("rbu_sa_reduce", c_void_p(x.ptr), x.ld, P, HW, C, p(a2g), p(a2g), p(a2g), p(ncarg), p(s_out), p(amax),
                                                       stream_ptr()), reps, flush), 1 * tb))
        pooled = ops.View(torch.empty((B, S // 2, S // 2, C), dtype=torch.bfloat16, device=dev))
        rows.append(("maxpool", timeit(lambda: call("rbu_maxpool2x2", c_void_p(x.ptr), x.ld, c_void_p(pooled.ptr), pooled.ld, B, S // 2,
                                                     S // 2, C, stream_ptr()), reps, flush), 1.25 * tb))
        for name, ms, gb in rows:
            print(f"{S:4d}x{S:<4d} C={C:<4d} {name:18s} {ms:7.3f} ms {gb / ms * 1e3:8.0f} GB/s  {gb / ms * 1e3 / PEAK:5.2f} of measured peak",
                  flush=True)


if __name__ == "__main__":
    main()
