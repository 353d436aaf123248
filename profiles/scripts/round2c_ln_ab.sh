#!/bin/bash
# LayerNorm A/B: parity of the product, then the layer_norm rows of profiles/gn_case.py for the product and under ";"-separated
# environment settings (FF_LN5_CTAS = grid cap of the sub-warp-row kernel)
timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x 2>&1 | tail -1
echo "== product"; python profiles/gn_case.py 2>&1 | grep layer_norm
IFS=';' read -ra EL <<< "$1"
for e in "${EL[@]}"; do echo "== $e"; env $e python profiles/gn_case.py 2>&1 | grep layer_norm; done
