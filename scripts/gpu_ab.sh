#!/bin/bash
# A/B of environment switches on the default bench: scripts/gpu_ab.sh NAME "ENV=.. ENV=.." [NAME "ENV.."]...
# every variant runs twice, interleaved (a fresh box's first run is slower), kernel table of the sparse convs printed.
set -u
mkdir -p gpurun_out
args=("$@")
for rep in 1 2; do
  i=0
  while [ $i -lt ${#args[@]} ]; do
    name=${args[$i]}; envs=${args[$((i+1))]}; i=$((i+2))
    env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ab_${name}_$rep.json 2> gpurun_out/ab_${name}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_${name}_$rep.json').read().strip().splitlines()[-1])
    ks={k['kernel']:k for k in d['kernels']}
    sel=[k for k in d['kernels'] if k['ms_per_step']>0.15]
    print('$name rep$rep', round(d['value'],1), 'sc/s', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], ' | '.join('%s %.3f'%(k['kernel'].replace('spconv_tc','sc'), k['ms_per_step']) for k in sel))
except Exception as e:
    print('$name rep$rep FAILED', e)
PY
  done
done
