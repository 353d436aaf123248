#!/bin/bash
# Round-1 extras (one B200): whole GPU suite, 768^2 line (config 5 shape: S=9216/2304 attention), channels_last A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/x_pytest.txt 2>&1; tail -3 gpurun_out/x_pytest.txt
timeout 100 python profiles/attn_case.py 5 > gpurun_out/x_attn_case.txt 2>&1; cat gpurun_out/x_attn_case.txt
timeout 600 python bench.py --res 768 --start-step 15 --edits 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/x_bench_768.json 2> gpurun_out/x_bench_768.err; tail -c 900 gpurun_out/x_bench_768.json; tail -2 gpurun_out/x_bench_768.err
timeout 600 python bench.py --start-step 35 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --channels-last > gpurun_out/x_bench_cl.json 2> gpurun_out/x_bench_cl.err; python -c "
import json; d=json.load(open('gpurun_out/x_bench_cl.json')); print('channels_last', d['value'], d['ms_per_step'])"
