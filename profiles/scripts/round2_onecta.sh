#!/bin/bash
# Round-2 starting point: first GPU run of the experimental one-CTA attention variant (csrc/attn_onecta.cuh).
# Build the variant HERE first (it must travel with the snapshot):
#   python -m freefine_b200.csrc.build --variant=onecta -DFF_ONE_CTA
# then:  gpurun --timeout 600 -- 'bash profiles/scripts/round2_onecta.sh'
mkdir -p gpurun_out
V=$PWD/freefine_b200/lib/libfreefine_b200_onecta.so
test -f "$V" || { echo "variant library missing: build it before gpurun"; exit 1; }
# parity first (a protocol bug shows up as a loud mbarrier time-out or a mismatch), under a short timeout
FREEFINE_B200_LIB=$V timeout 300 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > gpurun_out/oc_pytest.txt 2>&1; tail -5 gpurun_out/oc_pytest.txt
# timing of the dominant launch: product kernel, then the variant
timeout 100 python profiles/attn_case.py 5 > gpurun_out/oc_attn_case_product.txt 2>&1; cat gpurun_out/oc_attn_case_product.txt
FREEFINE_B200_LIB=$V timeout 100 python profiles/attn_case.py 5 > gpurun_out/oc_attn_case_variant.txt 2>&1; cat gpurun_out/oc_attn_case_variant.txt
