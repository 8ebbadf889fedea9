"""Per-kernel DRAM traffic and tensor-pipe activity from an `ncu --set full` capture of one training step
(`ncu -i x.ncu-rep --page raw --csv > x.csv`): writes profiles/traffic.json (read by bench.py for roofline.traffic) and
prints a markdown table.  Usage: python tools/ncu_traffic.py raw.csv "source description" > profiles/rNN_ncu_gemm.md"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS = {"conv_halo_kernel": "conv3x3", "conv_gemm_kernel": "conv_small", "wgrad_halo_kernel": "wgrad_gemm",
         "wgrad_gemm_kernel": "wgrad_gemm"}


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    source = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def scaled(r, name):          # bytes metrics come with a unit row (Kbyte / Mbyte / Gbyte)
        v, u = num(r[ix[name]]), units[ix[name]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    agg = collections.OrderedDict()
    print(f"# {source}\n")
    print("| kernel | grid | us | dram read MB | dram write MB | tensor pipe active % | issue active % | L2 throughput % |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        short = next((k for k in CLASS if k in name), None)
        if short is None:
            continue
        rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
        us = num(r[ix["gpu__time_duration.sum"]])
        tp = num(r[ix["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])
        ia = num(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]])
        l2 = num(r[ix["lts__throughput.avg.pct_of_peak_sustained_elapsed"]])
        print(f"| {short} | {r[ix['launch__grid_size']]} | {us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {tp:.1f} | {ia:.1f} | {l2:.1f} |")
        a = agg.setdefault(CLASS[short], {"launches": 0, "bytes": 0.0, "us": 0.0, "tp_us": 0.0})
        a["launches"] += 1
        a["bytes"] += rd + wr
        a["us"] += us
        a["tp_us"] += tp * us
    out = {}
    print("\n| class | launches | DRAM bytes / launch (MB) | time-weighted tensor pipe active % |\n|---|---|---|---|")
    for k, a in agg.items():
        out[k] = {"dram_bytes_per_launch": a["bytes"] / a["launches"], "launches": a["launches"], "source": source,
                  "tensor_pipe_active_pct": a["tp_us"] / a["us"] if a["us"] else None}
        print(f"| {k} | {a['launches']} | {a['bytes'] / a['launches'] / 1e6:.1f} | {a['tp_us'] / a['us']:.1f} |")
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
