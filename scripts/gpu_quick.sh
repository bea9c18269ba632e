#!/bin/bash
# Short GPU pass: targeted parity tests, bench, one ncu --set full capture of the sparse-conv launches.
set -u
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_features.py -m gpu -q --timeout 200 -x --no-header -k "spconv or bias_act_sum or dense_cnn" 2>&1 | tail -40 > gpurun_out/pytest_spconv.log
rc=${PIPESTATUS[0]}
echo "pytest spconv exit: $rc" >> gpurun_out/pytest_spconv.log
tail -5 gpurun_out/pytest_spconv.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 400 ncu --set full --clock-control none -k regex:k_spconv_tn -c 18 -o gpurun_out/ncu_spconv_tn \
  python bench.py --no-graph --batch 32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_tn.log 2>&1
echo "ncu full exit: $?" >> gpurun_out/ncu_tn.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -x --no-header 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
