mkdir -p gpurun_out
timeout 200 python tests/gpu_diag_attn.py > gpurun_out/r59_diag_f16.txt 2>&1; grep "max err\|FAILED\|Error\|timed out" gpurun_out/r59_diag_f16.txt | head -30
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x > gpurun_out/r59_pytest.txt 2>&1; tail -5 gpurun_out/r59_pytest.txt
for v in "" _p25; do
  echo "== variant '$v'" >> gpurun_out/r59_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r59_attn_case.txt 2>&1
done
FF_P=bf16x2 timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r59_attn_case.txt 2>&1
cat gpurun_out/r59_attn_case.txt
