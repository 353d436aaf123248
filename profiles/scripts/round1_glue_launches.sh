#!/bin/bash
# ncu launch list of one inversion + one sampling UNet call with the channels-last fast path (shares of the step).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --start-step 49 --no-cpu-baseline --no-e2e > gpurun_out/q_bench_under_ncu.log 2>&1
tail -2 gpurun_out/q_bench_under_ncu.log | cut -c1-300
wc -l gpurun_out/q_launches_bench.csv
