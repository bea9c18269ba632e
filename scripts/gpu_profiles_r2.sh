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
# the 20 k_spconv_tn launches of a forward come in layer order (6 x Cin 16, 5 x Cin 32, then 64->64 x 4, 64->128,
# 128->128 x 4): skip three warm-up forwards + 11 launches -> the four 64->64 launches, then the four 128->128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spconv_tn -s 71 -c 9 -o gpurun_out/ncu_spconv_tn64_r2 $B > gpurun_out/ncu_tn.log 2>&1
echo "tn64 $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -s 261 -c 12 -o gpurun_out/ncu_linear_r2 $B > gpurun_out/ncu_lin.log 2>&1
echo "lin $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mha_tc2 -s 9 -c 3 -o gpurun_out/ncu_mha_r2 $B > gpurun_out/ncu_mha.log 2>&1
echo "mha $?"
ls -la gpurun_out/*.ncu-rep
