#!/bin/bash
# Experiment pass: parity tests of the changed kernels (tile sort, known answers, model), bench A/B.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --timeout 300 -x --no-header -k "spconv or sort_tiles or known_answer or dense or forward" 2>&1 | tail -40 > gpurun_out/pytest_spconv.log
rc=${PIPESTATUS[0]}
echo "pytest exit: $rc" >> gpurun_out/pytest_spconv.log
tail -6 gpurun_out/pytest_spconv.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
U3D_SORT_TILES=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosort.json 2> gpurun_out/bench_nosort.err
cut -c1-300 gpurun_out/bench_nosort.json
tail -3 gpurun_out/bench.err
