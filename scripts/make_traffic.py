#!/usr/bin/env python
"""profiles/traffic.json from the per-kernel ncu pass (scripts/gpu_profiles_r2.sh -> ncu_kernels.csv): DRAM bytes per
launch (dram__bytes_read + write) and tensor-pipe activity of every sparse-conv layer shape, the linear kernel and the
attention kernel, keyed like bench.py's kernel table. The launches of one batch-32 SUN-RGBD forward come in the encoder's
layer order, which is how a launch is attributed to a layer shape. Records the commit of the captured binary.
usage: make_traffic.py gpurun_out/ncu_kernels.csv <commit> > profiles/traffic.json"""
import csv
import json
import sys

ENCODER_ORDER = ["16->16"] * 5 + ["16->32"] + ["32->32"] * 4 + ["32->64"] + ["64->64"] * 4 + ["64->128"] + ["128->128"] * 4


def main():
    path, commit = sys.argv[1], sys.argv[2]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = {}
    for r in csv.DictReader(lines):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * unit
        except ValueError:
            pass
    out = {}

    def add(key, d, kernel):
        a = out.setdefault(key, dict(dram=0.0, tp=0.0, n=0, kernel=kernel))
        a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["n"] += 1
    tn = [d for i, d in sorted(launches.items()) if "k_spconv_tn" in d["name"]]
    per_fwd = len(ENCODER_ORDER)
    for i, d in enumerate(tn[:per_fwd * (len(tn) // per_fwd)]):
        add(f"spconv_tc[27x{ENCODER_ORDER[i % per_fwd]}]", d, d["name"].split("(")[0].replace("void ", ""))
    for i, d in sorted(launches.items()):
        if "k_linear_tc" in d["name"]:
            add("linear_tc[all shapes]", d, "lin::k_linear_tc")
        elif "k_mha_tc2" in d["name"]:
            add("mha_core", d, "mha2::k_mha_tc2")
        elif "k_spconv_tc<" in d["name"]:
            add("spconv_tc[1x128->256]", d, "tc::k_spconv_tc")
    res = {}
    for key, a in out.items():
        res[key] = {"dram_bytes_per_launch": a["dram"] / a["n"], "launches_captured": a["n"], "batch": 32,
                    "tensor_pipe_active_pct": a["tp"] / a["n"], "kernel": a["kernel"], "commit": commit,
                    "source": "profiles/r02_ncu_kernels.csv (ncu --clock-control none, bench.py --no-graph --batch 32; "
                              "scripts/gpu_profiles_r2.sh)"}
    json.dump(res, sys.stdout, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
