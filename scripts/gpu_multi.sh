#!/bin/bash
# N-GPU pass (gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'): inference bench, training bench with the NCCL log.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?" >> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err; cut -c1-260 gpurun_out/bench_n$N.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --train --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_train$N.json 2> gpurun_out/bench_train$N.err
echo "exit $?" >> gpurun_out/bench_train$N.err
grep -c "AllReduce" gpurun_out/bench_train$N.err; tail -1 gpurun_out/bench_train$N.json | cut -c1-300
