"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST step
(launches after the last `stem_im2col_kernel`, which starts a step) and the individual GEMM launches."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:70]


def main():
    path = sys.argv[1]
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    starts = [i for i, r in enumerate(rows) if "stem_im2col" in r["Kernel Name"]]
    step = rows[starts[-1]:] if starts else rows
    tot = sum(float(r["Metric Value"]) for r in step) / 1e6
    agg = collections.OrderedDict()
    for r in step:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e6
    print(f"# last step: {len(step)} launches, {tot:.2f} ms serialised (cold-cache, under ncu)\n")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    print("\n## GEMM launches in order (grid, ms)\n")
    for r in step:
        if "gemm_kernel" in r["Kernel Name"]:
            print(f"{short(r['Kernel Name'])} grid {r['Grid Size']} {float(r['Metric Value']) / 1e6:.4f} ms")


if __name__ == "__main__":
    main()
