#!/bin/bash
# Round 2 (second session): attention A/B on one GPU box -- parity of the product build, then CUDA-event timing of the
# experiment variants built HERE beforehand (python -m freefine_b200.csrc.build --variant=<v> -D...).
# usage: round2b_attn_ab.sh <tag> "<variants>" "<ring modes for the product lib>"
mkdir -p gpurun_out
T=${1:-ab}; VARS=${2:-""}; MODES=${3:-""}
timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_kernels.py -q -m gpu -x --timeout 60 > gpurun_out/${T}_pytest.txt 2>&1; echo "product: $(tail -1 gpurun_out/${T}_pytest.txt)"
echo "== product" > gpurun_out/${T}_attn_case.txt
timeout 120 python profiles/attn_case.py 5 >> gpurun_out/${T}_attn_case.txt 2>&1
for v in $VARS; do
  echo "== variant $v" >> gpurun_out/${T}_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/${T}_attn_case.txt 2>&1
done
for m in $MODES; do
  echo "== product, FF_ATTN_RING=$m" >> gpurun_out/${T}_attn_case.txt
  FF_ATTN_RING=$m timeout 120 python profiles/attn_case.py 5 >> gpurun_out/${T}_attn_case.txt 2>&1
  FF_ATTN_RING=$m timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > gpurun_out/${T}_pytest_ring$m.txt 2>&1; echo "ring mode $m: $(tail -1 gpurun_out/${T}_pytest_ring$m.txt)"
done
cat gpurun_out/${T}_attn_case.txt
