#!/bin/bash
# Round-1 evidence run (one B200): ncu launch list of a short bench, ncu --set full of the dominant kernel launched in
# isolation (profiles/attn_case.py) and of the warp/blend kernel, CUDA-event timings outside ncu, HBM roofline script.
mkdir -p gpurun_out
timeout 100 python profiles/attn_case.py 5 > gpurun_out/p_attn_case_timing.txt 2>&1
timeout 200 python profiles/hbm_kernels.py > gpurun_out/p_hbm_kernels.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --start-step 49 --no-cpu-baseline --no-e2e > gpurun_out/p_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_masked_kv -c 1 -o gpurun_out/p_attn_full -f \
    python profiles/attn_case.py 1 > gpurun_out/p_ncu_attn.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:warp_affine -c 1 -s 2 -o gpurun_out/p_warp_full -f \
    python profiles/warp_case.py > gpurun_out/p_ncu_warp.log 2>&1
ls -la gpurun_out | tail -8
