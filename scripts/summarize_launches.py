#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv [steps_in_capture] > profiles/<name>.md"""
import csv
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg, tot = {}, 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
        a = agg.setdefault(r["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    ours = sum(v for n, (c, v) in agg.items() if any(t in n for t in ("u3d::", "tc::k_", "tn::k_", "mha::k_", "mha2::k_", "lin::k_", "k_linear_pack")))
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches captured (~{steps:g} steps), total {tot:.3f} ms of kernel time "
          f"(cold-cache, serialised: compare shares, not absolutes); libu3d_b200 kernels {ours:.3f} ms "
          f"({100 * ours / tot:.1f} %).\n")
    print("| ms total | share | launches | ms/launch | kernel |\n|---:|---:|---:|---:|---|")
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
        print(f"| {v:.3f} | {100 * v / tot:.1f} % | {c} | {v / c:.4f} | `{n[:100]}` |")


if __name__ == "__main__":
    main()
