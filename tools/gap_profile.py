"""Diagnostic (GPU): kernel timeline of one training step from torch.profiler (CUPTI) -- how much of the step the GPU
spends BETWEEN kernels of the compute stream (launch latency / dependency gaps) versus inside them."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rbunet                                     # noqa: E402
from tools.synthetic import synthetic_batch       # noqa: E402


def main():
    B, S = 64, 256
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = rbunet.RobustUNet(3, 1, 64).to(dev).train()
    crit = rbunet.RobustBCEDiceLoss()
    opt = rbunet.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    x, y = synthetic_batch(B, 3, S, S, seed=123)
    x, y = x.to(dev), y.to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):          # no synchronisation in between: the host runs ahead as in the benchmark loop
            step()
        torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", "trace.json")
    prof.export_chrome_trace(path)
    allev = json.load(open(path))["traceEvents"]
    ev = [e for e in allev if e.get("cat") == "kernel"]
    packs = sorted(e["ts"] + e["dur"] for e in ev if "pack_multi" in e["name"])
    stems = sorted(e["ts"] for e in ev if "stem_im2col" in e["name"])
    if len(packs) >= 2 and len(stems) >= 2:
        lo, hi = packs[1] - 500, stems[1] + 50
        print(f"events of any category between the 2nd pack kernel and the 2nd stem kernel ({(stems[1] - packs[1]):.0f} us):")
        for e in sorted((e for e in allev if "ts" in e and lo <= e["ts"] <= hi and e.get("ph") == "X"), key=lambda e: e["ts"]):
            print(f"   {e['ts'] - packs[1]:9.1f} us  dur {e.get('dur', 0):9.1f}  {e.get('cat', ''):14s} {e['name'][:70]}  stream {e.get('args', {}).get('stream')}")
    streams = {}
    for e in ev:
        streams.setdefault(e["args"].get("stream"), []).append((e["ts"], e["ts"] + e["dur"], e["name"]))
    for sid, ks in sorted(streams.items(), key=lambda kv: -len(kv[1])):
        ks.sort()
        big = sorted(((ks[i + 1][0] - ks[i][1], ks[i][2][:50], ks[i + 1][2][:50]) for i in range(len(ks) - 1)), reverse=True)[:8]
        for g, a, b in big:
            print(f"   gap {g:8.1f} us between {a} -> {b}")
        busy = sum(b - a for a, b, _ in ks)
        gaps = [max(0.0, ks[i + 1][0] - ks[i][1]) for i in range(len(ks) - 1)]
        span = ks[-1][1] - ks[0][0]
        small = sorted(g for g in gaps if g < 20)
        print(f"stream {sid}: {len(ks)} kernels, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, gaps {sum(gaps) / 1e3:.2f} ms "
              f"(gaps < 20 us: n={len(small)}, sum {sum(small) / 1e3:.2f} ms, median {small[len(small) // 2] if small else 0:.2f} us); "
              f"largest gaps: {[round(g, 1) for g in sorted(gaps)[-6:]]}")
    import collections
    import re
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        nm = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")
        nm = re.sub(r"\(.*", "", nm)[:40]
        agg[nm][0] += 1
        agg[nm][1] += e["dur"]
    print("per-kernel busy time per step (us), 3 steps averaged:")
    for nm, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f"   {nm:40s} {n // 3:4d} launches {us / 3:9.1f} us")
    os.remove(path)


if __name__ == "__main__":
    main()
