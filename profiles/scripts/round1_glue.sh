#!/bin/bash
# UNet-body glue kernels (csrc/unet_glue.cu): parity, the whole GPU suite, A/B bench at the GeoBench-2D schedule, rooflines.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x -s > gpurun_out/g_pytest_glue.txt 2>&1; tail -15 gpurun_out/g_pytest_glue.txt
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/g_pytest.txt 2>&1; tail -5 gpurun_out/g_pytest.txt
timeout 300 python bench.py --start-step 35 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_ss35_fast.json 2> gpurun_out/g_bench_fast.err; tail -c 1500 gpurun_out/g_bench_ss35_fast.json; tail -5 gpurun_out/g_bench_fast.err
timeout 300 python bench.py --start-step 35 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --plain-unet > gpurun_out/g_bench_ss35_plain.json 2> gpurun_out/g_bench_plain.err; tail -c 400 gpurun_out/g_bench_ss35_plain.json
timeout 200 python profiles/hbm_kernels.py > gpurun_out/g_hbm.json 2> gpurun_out/g_hbm.err; python -c "
import json; d=json.load(open('gpurun_out/g_hbm.json')); [print(k, round(v['gbs']), round(v['frac'],3), v.get('eager_ms_best'), v['ms_best']) for k,v in d.items() if isinstance(v,dict)]"; tail -3 gpurun_out/g_hbm.err
