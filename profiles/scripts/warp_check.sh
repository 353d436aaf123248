mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/w_pytest.txt 2>&1; tail -3 gpurun_out/w_pytest.txt
timeout 200 python profiles/hbm_kernels.py > gpurun_out/w_hbm.json 2>&1; cat gpurun_out/w_hbm.json | python -c "import json,sys; d=json.load(sys.stdin); [print(k, round(v['gbs']), round(v['frac'],3)) for k,v in d.items() if isinstance(v,dict)]"
