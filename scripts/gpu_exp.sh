#!/bin/bash
# Experiment pass: parity tests of the changed kernels, bench, A/B of the attention launch shape.
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_features.py -m gpu -q --timeout 200 -x --no-header -k "spconv or mha_core or head_forward" 2>&1 | tail -40 > gpurun_out/pytest_spconv.log
rc=${PIPESTATUS[0]}
echo "pytest exit: $rc" >> gpurun_out/pytest_spconv.log
tail -5 gpurun_out/pytest_spconv.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
U3D_MHA_QSPLIT=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qsplit.json 2> gpurun_out/bench_qsplit.err
cut -c1-300 gpurun_out/bench_qsplit.json
