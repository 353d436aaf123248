mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r51_pytest.txt 2>&1; tail -6 gpurun_out/r51_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r51_smoke.txt 2>&1; tail -2 gpurun_out/r51_smoke.txt
timeout 600 python bench.py --steps 1 --warmup 1 --start-step 35 --no-cpu-baseline --no-e2e > gpurun_out/r51_bench_ss35.json 2> gpurun_out/r51_bench.err; cat gpurun_out/r51_bench_ss35.json; tail -3 gpurun_out/r51_bench.err
