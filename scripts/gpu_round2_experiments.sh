#!/bin/bash
# First GPU call of round 2: run the gated tests of the code written after round 1's GPU budget was spent,
# then A/B every experimental switch on the default bench. ~4 GPU-minutes.
#   gpurun --timeout 600 -- 'bash scripts/gpu_round2_experiments.sh'
set -u
mkdir -p gpurun_out
U3D_EXPERIMENTAL=1 timeout 400 python -m pytest tests -m gpu -q --timeout 200 --no-header \
  -k "grouped or v_mn_major or points_prepare or points_gather" 2>&1 | tail -40 > gpurun_out/pytest_experimental.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_experimental.log
tail -8 gpurun_out/pytest_experimental.log
# the 4-stage sparse conv has no test of its own: run the conv parity tests under the switch
U3D_TN_SLICE_BUFS=1 timeout 300 python -m pytest tests/test_gpu_features.py tests/test_known_answers.py -m gpu -q --timeout 200 \
  --no-header -k "spconv or sparse_conv" 2>&1 | tail -8 > gpurun_out/pytest_tn_deep.log
tail -3 gpurun_out/pytest_tn_deep.log
run() {   # name, env assignments...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline --no-extra > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "$name: $(cut -c1-110 gpurun_out/bench_$name.json)"
}
run default U3D_NOP=1
run tn_deep U3D_TN_SLICE_BUFS=1
run mha_vmn U3D_MHA_VMN=1
run sort_group1_cin128 U3D_SORT_GROUP=1 U3D_SORT_MAX_CIN=128
run sort_group4_cin128 U3D_SORT_GROUP=4 U3D_SORT_MAX_CIN=128
run sort_group4 U3D_SORT_GROUP=4
run all U3D_TN_SLICE_BUFS=1 U3D_MHA_VMN=1 U3D_SORT_GROUP=4 U3D_SORT_MAX_CIN=128
