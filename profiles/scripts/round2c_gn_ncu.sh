#!/bin/bash
# ncu --set full of the GroupNorm kernels at 32x320x64x64: two-kernel form (FF_GN_CLUSTER=0) and the cluster form.
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.per_cycle_active,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread,launch__occupancy_limit_registers,launch__waves_per_multiprocessor,sm__cycles_elapsed.max"
FF_GN_CLUSTER=0 timeout 300 ncu --set full --clock-control none -k regex:gn_ -s 4 -c 2 -o gpurun_out/r2c_gn_two -f python profiles/gn_one.py 32 320 64 64 3 > gpurun_out/r2c_gn_two.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gn_ -s 2 -c 1 -o gpurun_out/r2c_gn_cluster -f python profiles/gn_one.py 32 320 64 64 3 > gpurun_out/r2c_gn_cluster.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gn_ -s 2 -c 1 -o gpurun_out/r2c_gn_small -f python profiles/gn_one.py 32 640 32 32 3 > gpurun_out/r2c_gn_small.log 2>&1
for f in r2c_gn_two r2c_gn_cluster r2c_gn_small; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ls -la gpurun_out/$f.ncu-rep
done
tail -3 gpurun_out/r2c_gn_two.log
