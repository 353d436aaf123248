#!/bin/bash
# Second pass over the glue kernels (parallel statistics reductions, MUFU SiLU / GELU, multi-row LayerNorm).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x > gpurun_out/h_pytest_glue.txt 2>&1; tail -4 gpurun_out/h_pytest_glue.txt
timeout 200 python profiles/hbm_kernels.py > gpurun_out/h_hbm.json 2> gpurun_out/h_hbm.err; python -c "
import json; d=json.load(open('gpurun_out/h_hbm.json')); [print(k, round(v['gbs']), round(v['frac'],3), v.get('eager_ms_best'), v['ms_best']) for k,v in d.items() if isinstance(v,dict)]"; tail -3 gpurun_out/h_hbm.err
timeout 300 python bench.py --start-step 35 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/h_bench_ss35.json 2> gpurun_out/h_bench.err; cut -c1-330 gpurun_out/h_bench_ss35.json; tail -3 gpurun_out/h_bench.err
