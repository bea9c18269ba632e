#!/usr/bin/env python
"""Markdown summary of an `ncu --set full --import-source on` capture: per launch the key raw metrics (duration, DRAM
bytes, L2 -> SM bytes over the crossbar, L2 / L1 throughput, hit rate, tensor-pipe activity, issue-slot use) and, for
the first launch, the SASS lines with the most stall samples.
usage: summarize_ncu_full.py capture.ncu-rep "<title>" > profiles/xxx.md   (runs `ncu -i ... --page raw/source --csv`)"""
import csv
import io
import subprocess
import sys

RAW = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
       ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"),
       ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 thr %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
       ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
       ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
       ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"), ("launch__registers_per_thread", "regs"),
       ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, title = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), lab, units[hdr.index(m)]) for m, lab in RAW if m in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# {title}\n\nSource: `{rep}` (`ncu --set full --clock-control none --import-source on`, bench.py --no-graph --batch 32; "
          "times under ncu are serialised, cold-cache).\n")
    print("| launch | kernel | " + " | ".join(f"{lab} ({u})" if u and u != "%" else lab for _, lab, u in cols) + " |")
    print("|---:|---|" + "---:|" * len(cols))
    for i, r in enumerate(body):
        name = r[kn].split("(")[0].replace("void ", "")
        vals = []
        for c, lab, u in cols:
            try:
                v = float(r[c].replace(",", ""))
                vals.append(f"{v:.1f}" if v < 1e4 else f"{v:.0f}")
            except ValueError:
                vals.append(r[c])
        print(f"| {i} | `{name}` | " + " | ".join(vals) + " |")
    # derived: L2 -> SM rate of each launch
    try:
        ti, bi = hdr.index("gpu__time_duration.sum"), hdr.index("l1tex__m_xbar2l1tex_read_bytes.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}
        print("\nL2 -> SM rate (crossbar read bytes / duration): " + ", ".join(
            f"{float(r[bi]) * scale.get(units[bi], 1.0) / (float(r[ti]) * tsc.get(units[ti], 1e-6)) / 1e12:.2f} TB/s" for r in body))
    except ValueError:
        pass
    src = ncu_csv(rep, "source")
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not heads:
        return
    h = src[heads[0]]
    blk = src[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(src))]
    si, so = h.index("# Samples"), h.index("Source")
    samp = [(int(r[si]), i) for i, r in enumerate(blk) if len(r) > si and r[si].isdigit()]
    tot = sum(s for s, _ in samp) or 1
    print(f"\n## Launch 0: SASS lines with the most warp-stall samples ({tot} samples)\n")
    print("| line | samples | share | instruction |\n|---:|---:|---:|---|")
    for s, i in sorted(sorted(samp, reverse=True)[:18], key=lambda t: t[1]):
        print(f"| {i} | {s} | {100.0 * s / tot:.1f} % | `{blk[i][so].strip()[:90]}` |")


if __name__ == "__main__":
    main()
