#!/bin/bash
# gpurun_out/ of scripts/gpu_r2_full.sh -> profiles/r02_* (run here, after the GPU pass): collect_profiles.sh <commit>
set -u
C=${1:-$(git rev-parse --short HEAD)}
cp gpurun_out/bench_final.json profiles/r02_bench_final.json
cp gpurun_out/bench_ref.json profiles/r02_bench_ref.json
(tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/gpu.txt) > profiles/r02_gpu_tests.txt
cp gpurun_out/launches_b32.csv profiles/r02_launches_b32.csv
python scripts/summarize_launches.py gpurun_out/launches_b32.csv > profiles/r02_launches_b32.md
cp gpurun_out/ncu_kernels.csv profiles/r02_ncu_kernels.csv
python scripts/summarize_ncu_kernels.py gpurun_out/ncu_kernels.csv > profiles/r02_ncu_kernels.md
python scripts/make_traffic.py gpurun_out/ncu_kernels.csv $C > profiles/traffic.json
python scripts/summarize_ncu_full.py gpurun_out/ncu_spconv_tn64_r2.ncu-rep "k_spconv_tn: the four 64->64 launches (tcgen05.mma.ws, M = 64), 64->128 and the four 128->128 launches of one forward" > profiles/r02_ncu_spconv_tn64.md
python scripts/summarize_ncu_full.py gpurun_out/ncu_linear_r2.ncu-rep "k_linear_tc: twelve launches of the decoder" > profiles/r02_ncu_linear.md
python scripts/summarize_ncu_full.py gpurun_out/ncu_mha_r2.ncu-rep "k_mha_tc2: the three self-attention launches" > profiles/r02_ncu_mha.md
[ -s gpurun_out/bench_linear.txt ] && cp gpurun_out/bench_linear.txt profiles/r02_linear_ab.txt
[ -s gpurun_out/bench_mha.txt ] && cp gpurun_out/bench_mha.txt profiles/r02_mha_ab.txt
