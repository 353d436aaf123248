#!/bin/bash
# Final validation of the committed tree (one B200): GPU suite, smoke, HBM roofline script.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/f_pytest.txt 2>&1; tail -3 gpurun_out/f_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.txt 2>&1; tail -1 gpurun_out/f_smoke.txt
timeout 200 python profiles/hbm_kernels.py > gpurun_out/f_hbm.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/f_hbm.json')); [print(k, round(v['gbs']), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)]"
