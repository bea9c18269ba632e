#!/usr/bin/env python
"""Per-kernel table from `ncu --metrics ... --csv` (scripts/gpu_profiles_r2.sh): one row per kernel name with
mean time, DRAM bytes, achieved DRAM GB/s vs the measured copy peak, tensor-pipe activity, L2 hit rate.
usage: summarize_ncu_kernels.py ncu_kernels.csv [peak_GBs] > profiles/<name>.md"""
import csv
import json
import os
import sys


def main():
    path = sys.argv[1]
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if peak is None:
        pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6650.0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = {}
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        d["unit:" + r["Metric Name"]] = r["Metric Unit"]
    agg = {}
    for d in launches.values():
        name = d["name"].split("(")[0].replace("void ", "").replace("u3d::", "")
        a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, tp=0.0, l2=0.0, lts=0.0, regs=0))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d.get("unit:gpu__time_duration.sum", "ns")]
        bs = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        a["n"] += 1
        a["t"] += d.get("gpu__time_duration.sum", 0.0) * scale
        a["rd"] += d.get("dram__bytes_read.sum", 0.0) * bs.get(d.get("unit:dram__bytes_read.sum", "byte"), 1.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0) * bs.get(d.get("unit:dram__bytes_write.sum", "byte"), 1.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["l2"] += d.get("lts__t_sector_hit_rate.pct", 0.0)
        a["lts"] += d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["regs"] = int(d.get("launch__registers_per_thread", 0))
    tot = sum(a["t"] for a in agg.values())
    print(f"# Per-kernel ncu metrics of one forward (batch 32, bf16, eager launches): every libu3d kernel\n")
    print(f"Source: `{path}` (scripts/gpu_profiles_r2.sh). Times under ncu are serialised and cold-cache; DRAM GB/s = "
          f"(dram read + write bytes) / kernel time, % of the measured copy peak ({peak:.1f} GB/s, MEASURED_PEAKS.json); "
          f"tensor pipe = sm__pipe_tensor_cycles_active (% of active cycles); L2 = lts throughput % of peak.\n")
    print("| kernel | launches | total us | us/launch | DRAM MB/launch | DRAM GB/s | % of HBM peak | tensor pipe % | L2 throughput % | L2 hit % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        n = a["n"]
        gbs = (a["rd"] + a["wr"]) / max(a["t"], 1e-9) / 1e3
        print(f"| `{name[:48]}` | {n} | {a['t']:.1f} | {a['t'] / n:.1f} | {(a['rd'] + a['wr']) / n / 1e6:.2f} | {gbs:.0f} | "
              f"{100 * gbs / peak:.1f} | {a['tp'] / n:.1f} | {a['lts'] / n:.1f} | {a['l2'] / n:.0f} | {a['regs']} |")
    print(f"\nTotal libu3d kernel time of the capture under ncu: {tot / 1e3:.2f} ms over {sum(a['n'] for a in agg.values())} launches.")


if __name__ == "__main__":
    main()
