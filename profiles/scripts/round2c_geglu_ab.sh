#!/bin/bash
# GEGLU A/B: parity, then profiles/gn_case.py (GEGLU rows) for the product and under ";"-separated environment settings
# usage: round2c_geglu_ab.sh "<env settings>"
timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x -k "geglu or unet" 2>&1 | tail -1
echo "== product"; python profiles/gn_case.py 2>&1 | grep geglu
IFS=';' read -ra EL <<< "$1"
for e in "${EL[@]}"; do echo "== $e"; env $e python profiles/gn_case.py 2>&1 | grep geglu; done
