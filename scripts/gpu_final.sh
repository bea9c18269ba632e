#!/bin/bash
# Round-end GPU pass: full parity suite, smoke, bench (both arms), ncu launch list, per-kernel ncu metrics
# for every libu3d kernel of one forward, one --set full capture of the wide sparse-conv launches.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 400 --no-header 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-400 gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --no-graph --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-extra > gpurun_out/bench_ncu.log 2>&1
echo "ncu list exit: $?" >> gpurun_out/bench_ncu.log
timeout 400 ncu --clock-control none -k 'regex:k_' -c 220 --csv --log-file gpurun_out/ncu_kernels.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread \
  python bench.py --no-graph --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-extra > gpurun_out/ncu_kernels.log 2>&1
echo "ncu kernels exit: $?" >> gpurun_out/ncu_kernels.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tn -s 10 -c 8 -o gpurun_out/ncu_spconv_tn \
  python bench.py --no-graph --batch 32 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-extra > gpurun_out/ncu_tn.log 2>&1
echo "ncu full exit: $?" >> gpurun_out/ncu_tn.log
ls -la gpurun_out | tail -20
