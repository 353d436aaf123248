#!/bin/bash
# Final validation + evidence with the channels-last UNet fast path (one B200): GPU suite, smoke, the driver's default
# bench line (full 50+50 schedule), the GeoBench-2D schedule line, graph-timed HBM rooflines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/k_pytest.txt 2>&1; tail -3 gpurun_out/k_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.txt 2>&1; tail -1 gpurun_out/k_smoke.txt | cut -c1-200
timeout 300 python profiles/hbm_kernels.py > gpurun_out/k_hbm.json 2> gpurun_out/k_hbm.err; python -c "
import json; d=json.load(open('gpurun_out/k_hbm.json')); [print(k, round(v['gbs']), round(v['frac'],3), v.get('eager_ms_best'), round(v['ms_best'],4)) for k,v in d.items() if isinstance(v,dict)]"; tail -3 gpurun_out/k_hbm.err
timeout 300 python bench.py --start-step 35 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_ss35.json 2> gpurun_out/k_bench_ss35.err; cut -c1-300 gpurun_out/k_bench_ss35.json
timeout 900 python bench.py > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err; cut -c1-300 gpurun_out/k_bench_n1.json; tail -2 gpurun_out/k_bench_n1.err
