#!/bin/bash
# Last validation of the committed tree (one B200): GPU suite, smoke, GeoBench-2D-schedule bench (+ cuDNN autotune A/B), HBM rooflines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/z_pytest.txt 2>&1; tail -3 gpurun_out/z_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.txt 2>&1; tail -1 gpurun_out/z_smoke.txt | cut -c1-120
timeout 300 python bench.py --start-step 35 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/z_bench_ss35.json 2> gpurun_out/z_bench_ss35.err; cut -c1-200 gpurun_out/z_bench_ss35.json
timeout 300 python bench.py --start-step 35 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --cudnn-benchmark > gpurun_out/z_bench_ss35_cb.json 2> gpurun_out/z_bench_ss35_cb.err; cut -c1-200 gpurun_out/z_bench_ss35_cb.json; tail -2 gpurun_out/z_bench_ss35_cb.err
timeout 300 python profiles/hbm_kernels.py > gpurun_out/z_hbm.json 2> gpurun_out/z_hbm.err; python -c "
import json; d=json.load(open('gpurun_out/z_hbm.json')); [print(k, round(v['gbs']), round(v['frac'],3), round(v['ms_best'],4)) for k,v in d.items() if isinstance(v,dict) and ('norm' in k or 'geglu' in k or 'bias' in k)]"; tail -2 gpurun_out/z_hbm.err
