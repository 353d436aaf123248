#!/bin/bash
# LayerNorm A/B: parity of the product, then the layer_norm rows of profiles/gn_case.py for the product, under ";"-separated
# environment settings (FF_LN5_CTAS = grid cap of the sub-warp-row kernel) and for variant libraries
# usage: round2c_ln_ab.sh "<env settings>" "<variants>"
timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x 2>&1 | tail -1
echo "== product"; python profiles/gn_case.py 2>&1 | grep layer_norm
IFS=';' read -ra EL <<< "$1"
for e in "${EL[@]}"; do echo "== $e"; env $e python profiles/gn_case.py 2>&1 | grep layer_norm; done
for v in $2; do echo "== variant $v"; FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_$v.so python profiles/gn_case.py 2>&1 | grep layer_norm; done
