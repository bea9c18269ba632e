#!/bin/bash
# Round-2 evidence pass: full GPU parity suite, smoke, default bench (+ extra config blocks), reference arm, profiles.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --no-header 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench $?"
cut -c1-200 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-300 gpurun_out/bench_ref.json
timeout 200 python scripts/bench_linear.py > gpurun_out/bench_linear.txt 2>&1; head -3 gpurun_out/bench_linear.txt | cut -c1-120
timeout 200 python scripts/bench_mha.py > gpurun_out/bench_mha.txt 2>&1; head -4 gpurun_out/bench_mha.txt | cut -c1-120
bash scripts/gpu_profiles_r2.sh
