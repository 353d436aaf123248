#!/bin/bash
# Round 2: parity + timing + timeline of the ring attention kernel (csrc/attn_ring.cuh) in its layouts (FF_ATTN_RING modes).
# Build HERE first:  python -m freefine_b200.csrc.build && python -m freefine_b200.csrc.build --variant=tl -DFF_TIMELINE
# usage: round2_ring.sh <tag> "<parity modes>" "<timing modes>" "<timeline modes>"
mkdir -p gpurun_out
T=${1:-a}; PM=${2:-"3 4"}; TM=${3:-"0 1 2 3 4 5"}; LM=${4:-"3 4"}
for m in $PM; do
  FF_ATTN_RING=$m timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > gpurun_out/${T}_pytest_ring$m.txt 2>&1; echo "mode $m: $(tail -1 gpurun_out/${T}_pytest_ring$m.txt)"
done
for m in $TM; do
  FF_ATTN_RING=$m timeout 100 python profiles/attn_case.py 5 > gpurun_out/${T}_attn_case_ring$m.txt 2>&1; echo "mode $m"; cat gpurun_out/${T}_attn_case_ring$m.txt
done
V=$PWD/freefine_b200/lib/libfreefine_b200_tl.so
if test -f "$V"; then
  for m in $LM; do
    FREEFINE_B200_LIB=$V FF_ATTN_RING=$m timeout 100 python profiles/timeline_ring.py 4096 40 > gpurun_out/${T}_timeline_ring$m.txt 2>&1
    grep "tile period\|arrive" gpurun_out/${T}_timeline_ring$m.txt | head -12
  done
fi
