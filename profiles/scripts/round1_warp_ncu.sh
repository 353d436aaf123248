#!/bin/bash
# ncu --set full of the warp/blend kernel (fast path) on the L2-exceeding roofline batch.
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:warp_affine -c 1 -s 2 -o gpurun_out/p_warp_full -f \
    python profiles/warp_case.py > gpurun_out/p_ncu_warp.log 2>&1; tail -2 gpurun_out/p_ncu_warp.log
