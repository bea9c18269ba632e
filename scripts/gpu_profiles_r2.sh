#!/bin/bash
# Round-2 profile pass: ncu launch list of one batch-32 forward, per-kernel metrics of every libu3d kernel,
# --set full captures of the tcgen05 kernels that dominate (64->64 sparse conv, linear, attention).
set -u
mkdir -p gpurun_out
B="python bench.py --no-graph --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1600 --csv \
  --log-file gpurun_out/launches_b32.csv $B > gpurun_out/ncu_list.log 2>&1
echo "list $?"
timeout 600 ncu --clock-control none -k 'regex:k_' -s 620 -c 220 --csv --log-file gpurun_out/ncu_kernels.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread \
  $B > gpurun_out/ncu_kernels.log 2>&1
echo "kernels $?"
# the 64 -> 64 SubM layers run the <64, 2, 8> instantiation (M = 64 path): 4 launches per forward
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_spconv_tn<64, 2, 8' -s 12 -c 4 -o gpurun_out/ncu_spconv_tn64_r2 $B > gpurun_out/ncu_tn.log 2>&1
echo "tn64 $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -s 261 -c 12 -o gpurun_out/ncu_linear_r2 $B > gpurun_out/ncu_lin.log 2>&1
echo "lin $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mha_tc2 -s 9 -c 3 -o gpurun_out/ncu_mha_r2 $B > gpurun_out/ncu_mha.log 2>&1
echo "mha $?"
ls -la gpurun_out/*.ncu-rep
