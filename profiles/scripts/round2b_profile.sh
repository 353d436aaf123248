#!/bin/bash
# Round-2 (second session) evidence run (one B200): ncu launch list of a short bench (one inversion + one sampling UNet call of 8 edits),
# ncu --set full of the two product attention kernels launched in isolation (profiles/attn_case.py: legacy kernel for
# d = 40, ring kernel for d = 80), CUDA-event timings outside ncu.
mkdir -p gpurun_out
timeout 100 python profiles/attn_case.py 5 > gpurun_out/r2q_attn_case_timing.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q_launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --start-step 49 --no-cpu-baseline --no-e2e --no-extras --no-cudnn-benchmark > gpurun_out/r2q_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_masked_kv -c 1 -o gpurun_out/r2q_attn_d40_full -f \
    python profiles/attn_case.py 1 > gpurun_out/r2q_ncu_attn_d40.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_ring -c 1 -o gpurun_out/r2q_attn_d80_full -f \
    python profiles/attn_case.py 1 > gpurun_out/r2q_ncu_attn_d80.log 2>&1
ls -la gpurun_out | grep r2q_
