#!/bin/bash
# Single-read GroupNorm (csrc/unet_glue.cu gn_fused_nhwc_kernel): parity suite, then the HBM rooflines of the norm kernels
# with the single-read form and with the two-kernel form (FF_GN_TWO_KERNEL=1) side by side.
mkdir -p gpurun_out
FF_GN_SINGLE_READ=1 timeout 300 python -m pytest tests/test_gpu_unet_glue.py -q -m gpu -x 2>&1 | tail -3
cat > gpurun_out/gn_rf.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from freefine_b200 import roofline as RF
dev = torch.device("cuda:0")
rows = RF.hbm_rooflines(dev, 6550.1, with_eager=True)
tag = "single-read" if os.environ.get("FF_GN_SINGLE_READ") == "1" else "two-kernel"
for r in rows:
    if "norm" in r["kernel"]:
        print(tag, r["kernel"], r["size"], round(r["ms"] * 1e3, 1), "us frac", round(r["frac"], 3), "eager_us", round(r.get("eager_ms", 0) * 1e3, 1))
PY
FF_GN_SINGLE_READ=1 timeout 200 python gpurun_out/gn_rf.py 2>&1 | tail -5 | tee gpurun_out/gn_single_read.txt
FF_GN_SINGLE_READ=0 timeout 200 python gpurun_out/gn_rf.py 2>&1 | tail -5 | tee gpurun_out/gn_two_kernel.txt
for kb in 24 48 100; do echo "tile_kb=$kb"; FF_GN_SINGLE_READ=1 FF_GN_TILE_KB=$kb timeout 120 python profiles/gn_case.py 2>&1 | tail -6 | tee gpurun_out/gn_case_single_read_$kb.txt; done
echo two-kernel; timeout 120 python profiles/gn_case.py 2>&1 | tail -6 | tee gpurun_out/gn_case_two_kernel.txt
