mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu > gpurun_out/r60_pytest.txt 2>&1; tail -5 gpurun_out/r60_pytest.txt
for v in "" _p25 _p37 _p50 _ko; do
  echo "== variant '$v'" >> gpurun_out/r60_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r60_attn_case.txt 2>&1
done
cat gpurun_out/r60_attn_case.txt
FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_tl.so timeout 120 python profiles/timeline.py > gpurun_out/r60_timeline.txt 2>&1
tail -3 gpurun_out/r60_timeline.txt
