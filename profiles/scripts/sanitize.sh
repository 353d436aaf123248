#!/bin/bash
# compute-sanitizer over one small launch of every kernel (tests/gpu_kernel_tour.py): memcheck (out-of-bounds / misaligned
# global, shared and local accesses), then racecheck (shared-memory hazards), then the single-read and the thread-block-cluster GroupNorm experiments.
# Summaries -> gpurun_out/sanitize_*.txt (copied to profiles/r2_sanitize_*.txt).
#   gpurun --timeout 1200 -- 'bash profiles/scripts/sanitize.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_kernel_tour.py > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|kernel tour ok|Invalid|Error" gpurun_out/sanitize_memcheck.txt | head -8
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tests/gpu_kernel_tour.py > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|kernel tour ok|hazard|Error" gpurun_out/sanitize_racecheck.txt | head -8
FF_GN_SINGLE_READ=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_kernel_tour.py > gpurun_out/sanitize_memcheck_gn_single_read.txt 2>&1
echo "memcheck (single-read GroupNorm) rc=$?"; grep -E "ERROR SUMMARY|kernel tour ok" gpurun_out/sanitize_memcheck_gn_single_read.txt | head -3
FF_GN_CLUSTER=1 timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tests/gpu_kernel_tour.py > gpurun_out/sanitize_memcheck_gn_cluster.txt 2>&1
echo "memcheck (cluster GroupNorm) rc=$?"; grep -E "ERROR SUMMARY|kernel tour ok" gpurun_out/sanitize_memcheck_gn_cluster.txt | head -3
