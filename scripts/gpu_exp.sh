#!/bin/bash
# Experiment pass: which rulebooks to tile-sort (policy A/B of validated code paths).
set -u
mkdir -p gpurun_out
for cfg in "32 1" "32 0" "16 1"; do
  set -- $cfg
  U3D_SORT_MAX_CIN=$1 U3D_SORT_DOWN=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_sort_$1_$2.json 2> gpurun_out/bench_sort_$1_$2.err
  echo "cin<=$1 down=$2: $(cut -c1-120 gpurun_out/bench_sort_$1_$2.json)"
done
