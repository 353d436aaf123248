#!/bin/bash
# Round-1 bench lines (one B200): the driver's default invocation, the GeoBench-2D default schedule (start_step 35) and
# the reference arm.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; tail -c 600 gpurun_out/b_bench_n1.json
timeout 600 python bench.py --start-step 35 --steps 3 --warmup 3 > gpurun_out/b_bench_n1_startstep35.json 2> gpurun_out/b_bench_ss35.err; tail -c 300 gpurun_out/b_bench_n1_startstep35.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/b_bench_reference_arm.json 2> gpurun_out/b_bench_ref.err; tail -c 400 gpurun_out/b_bench_reference_arm.json
