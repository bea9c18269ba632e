#!/bin/bash
# Mid-round pass: MMA cost micro-benchmark, the sparse-conv switch test, default bench, profile pass.
set -u
mkdir -p gpurun_out
./scripts/ubench/mma_cost > gpurun_out/ubench_mma.txt 2>&1; echo "ubench $?"
timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -q -x --no-header -k "spconv" 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench $?"
cut -c1-200 gpurun_out/bench_final.json
bash scripts/gpu_profiles_r2.sh
