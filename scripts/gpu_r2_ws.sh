#!/bin/bash
# tcgen05.mma.ws experiment: layout + cost micro-benchmark, parity tests, A/B bench.
set -u
mkdir -p gpurun_out
./scripts/ubench/mma_cost > gpurun_out/ubench_mma.txt 2>&1; echo "ubench $?"
head -75 gpurun_out/ubench_mma.txt
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_geometry.py tests/test_known_answers.py -m gpu -q --no-header -k "spconv or fps" 2>&1 | tail -15
for ws in 0 1; do
  U3D_TN_WS=$ws timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_ws_$ws.json 2> gpurun_out/bench_ws_$ws.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ws_$ws.json').read().strip().splitlines()[-1])
print('ws=$ws', round(d['value'],1), d['ms_per_step'], d['clocks'])
for k in d['kernels']:
    if 'spconv' in k['kernel']: print('  ', k['kernel'], k['calls_per_step'], round(k['avg_launch_ms'],4), round(k['tflops'],1))
PY
done
