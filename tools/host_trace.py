"""Diagnostic (GPU): host-side timestamps of the first C-ABI calls of a steady-state training step (where does the
launching thread spend its time between `rbu_pack_weights_multi` and `rbu_stem_im2col`?)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet                                     # noqa: E402
from rbunet import engine as E                    # noqa: E402
from tools.synthetic import synthetic_batch       # noqa: E402

LOG = []
_orig_call = E.call


def traced(name, *a, **k):
    t0 = time.perf_counter()
    _orig_call(name, *a, **k)
    LOG.append((name, t0, time.perf_counter()))


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 64).to(dev).train()
    crit = rbunet.RobustBCEDiceLoss()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    x, y = synthetic_batch(64, 3, 256, 256, seed=123)
    x, y = x.to(dev), y.to(dev)
    E.call = traced
    marks = []
    for i in range(6):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        t1 = time.perf_counter()
        out = model(x)
        t2 = time.perf_counter()
        loss = crit(out, y)
        t3 = time.perf_counter()
        loss.backward()
        t4 = time.perf_counter()
        opt.step()
        t5 = time.perf_counter()
        marks.append((t0, t1, t2, t3, t4, t5))
        if i == 4:
            LOG.clear()
    torch.cuda.synchronize()
    t0, t1, t2, t3, t4, t5 = marks[-1]
    print(f"last step host times: zero_grad {1e3 * (t1 - t0):.2f} ms, forward {1e3 * (t2 - t1):.2f}, loss {1e3 * (t3 - t2):.2f}, "
          f"backward {1e3 * (t4 - t3):.2f}, opt.step {1e3 * (t5 - t4):.2f}")
    prev = t1
    for name, a, b in LOG[:12]:
        print(f"  +{1e3 * (a - prev):7.3f} ms  {name:28s} call took {1e3 * (b - a):.3f} ms")
        prev = b
    alloc = torch.cuda.memory_stats(dev)
    print("cudaMalloc calls so far:", alloc.get("num_device_alloc"), "cudaFree:", alloc.get("num_device_free"),
          "alloc retries:", alloc.get("num_alloc_retries"))


if __name__ == "__main__":
    main()
