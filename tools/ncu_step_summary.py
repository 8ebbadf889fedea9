"""Summarise gpurun_out/ncu_step.csv (tools/ncu_step.sh: long-format `ncu --csv` log, one row per launch and metric) for
ONE training step: per kernel name -> launches, total us, time-weighted DRAM throughput %, tensor-pipe %, issue %,
occupancy %, shared-memory operand wavefront %, DRAM bytes.  Usage: python tools/ncu_step_summary.py [csv] > profiles/..."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu_step.csv"
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    launches = collections.OrderedDict()
    for r in rows:
        lid = int(r["ID"])
        d = launches.setdefault(lid, {"name": r["Kernel Name"], "grid": r.get("Grid Size", ""), "block": r.get("Block Size", "")})
        v = r["Metric Value"].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        unit = r["Metric Unit"]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    ids = sorted(launches)
    # one step: from one stem_im2col launch to the next
    starts = [i for i in ids if "stem_im2col" in launches[i]["name"]]
    lo, hi = (starts[0], starts[1]) if len(starts) >= 2 else (ids[0], ids[-1] + 1)
    step = [launches[i] for i in ids if lo <= i < hi]
    T = "gpu__time_duration.sum"
    keys = {"dram%": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "occ%": "sm__warps_active.avg.pct_of_peak_sustained_active",
            "L2%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "smemTC%": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"}
    agg = collections.OrderedDict()
    for d in step:
        name = re.sub(r"\(.*", "", d["name"])
        name = re.sub(r"^void |\(anonymous namespace\)::", "", name)
        a = agg.setdefault(name, collections.defaultdict(float))
        us = d.get(T, 0.0)
        a["n"] += 1
        a["us"] += us
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["regs"] = d.get("launch__registers_per_thread", 0.0)
        for k, m in keys.items():
            a[k] += d.get(m, 0.0) * us
    total = sum(a["us"] for a in agg.values())
    print(f"# one training step (batch 64, 256x256): {len(step)} launches, {total / 1e3:.2f} ms serialised under ncu "
          f"(--clock-control none; cold-cache, compare shares)\n")
    print("| kernel | launches | us | share | DRAM % | tensor % | issue % | occupancy % | L2 % | smem->TC % | DRAM MB | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        us = a["us"] or 1.0
        print(f"| {name[:60]} | {int(a['n'])} | {a['us']:.0f} | {100 * a['us'] / total:.1f}% | " +
              " | ".join(f"{a[k] / us:.1f}" for k in keys) + f" | {a['bytes'] / 1e6:.0f} | {int(a['regs'])} |")
    if "--traffic" in sys.argv:
        # profiles/traffic.json: per GEMM class DRAM bytes per launch and time-weighted tensor-pipe activity (bench.py reads it
        # for roofline.traffic)
        import json
        import os
        source = sys.argv[sys.argv.index("--traffic") + 1]
        cls = {"conv_halo_pair_kernel": "conv3x3", "conv_halo_kernel": "conv3x3", "conv_gemm_kernel": "conv_small",
               "wgrad_halo_kernel": "wgrad_gemm", "wgrad_halo2_kernel": "wgrad_gemm", "wgrad_pair_kernel": "wgrad_gemm",
               "wgrad_gemm_kernel": "wgrad_gemm"}
        out = {}
        for name, a in agg.items():
            c = cls.get(name.replace("<unnamed>::", ""))
            if c is None:
                continue
            o = out.setdefault(c, {"launches": 0, "bytes": 0.0, "us": 0.0, "tp_us": 0.0})
            o["launches"] += int(a["n"])
            o["bytes"] += a["bytes"]
            o["us"] += a["us"]
            o["tp_us"] += a["tensor%"]
        res = {c: {"dram_bytes_per_launch": o["bytes"] / o["launches"], "launches": o["launches"], "source": source,
                   "tensor_pipe_active_pct": o["tp_us"] / o["us"]} for c, o in out.items()}
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        with open(os.path.join(root, "profiles", "traffic.json"), "w") as f:
            json.dump(res, f, indent=1)
        print("\n| class | launches | DRAM MB / launch | time-weighted tensor pipe % |\n|---|---|---|---|")
        for c, r in res.items():
            print(f"| {c} | {r['launches']} | {r['dram_bytes_per_launch'] / 1e6:.1f} | {r['tensor_pipe_active_pct']:.1f} |")
    if "--launches" in sys.argv:
        print("\n## launches in order\n")
        for d in step:
            name = re.sub(r"^void |\(anonymous namespace\)::|\(.*", "", d["name"])
            print(f"{name[:48]:48s} grid {d['grid']:>16s} {d.get(T, 0):8.1f} us  dram {d.get(keys['dram%'], 0):5.1f}%  "
                  f"tensor {d.get(keys['tensor%'], 0):5.1f}%  issue {d.get(keys['issue%'], 0):5.1f}%  occ {d.get(keys['occ%'], 0):5.1f}%")


if __name__ == "__main__":
    main()
