mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r65_pytest.txt 2>&1; tail -4 gpurun_out/r65_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r65_smoke.txt 2>&1; tail -2 gpurun_out/r65_smoke.txt
timeout 100 python profiles/attn_case.py 5 > gpurun_out/r65_attn_case.txt 2>&1; cat gpurun_out/r65_attn_case.txt
timeout 600 python bench.py --steps 1 --warmup 1 --start-step 35 --no-cpu-baseline --no-e2e > gpurun_out/r65_bench_ss35.json 2> gpurun_out/r65_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r65_bench_ss35.json')); print(d['value'], d['ms_per_step'], d['roofline'])"; tail -3 gpurun_out/r65_bench.err
