#!/bin/bash
# Round 2 (third session): GroupNorm A/B on one GPU box -- parity of the product build, then profiles/gn_case.py under a
# list of environment settings (";"-separated, e.g. "FF_GN_CLUSTER=0;FF_GN_CLUSTER=0 FF_GN_CHUNK_PX=111") and for
# variant libraries built beforehand (python -m freefine_b200.csrc.build --variant=<v> -D...).
# usage: round2c_gn_ab.sh <tag> "<env settings>" "<variants>"
mkdir -p gpurun_out
T=${1:-gn}; ENVS=${2:-""}; VARS=${3:-""}
timeout 600 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x --timeout 120 > gpurun_out/${T}_pytest.txt 2>&1; echo "product: $(tail -1 gpurun_out/${T}_pytest.txt)"
O=gpurun_out/${T}_gn_case.txt
echo "== product" > $O; timeout 200 python profiles/gn_case.py >> $O 2>&1
IFS=';' read -ra EL <<< "$ENVS"
for e in "${EL[@]}"; do
  echo "== product $e" >> $O; env $e timeout 200 python profiles/gn_case.py >> $O 2>&1
done
for v in $VARS; do
  echo "== variant $v" >> $O
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_$v.so timeout 200 python profiles/gn_case.py >> $O 2>&1
  for e in "${EL[@]}"; do
    echo "== variant $v $e" >> $O; env $e FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_$v.so timeout 200 python profiles/gn_case.py >> $O 2>&1
  done
done
grep -v layer_norm $O
