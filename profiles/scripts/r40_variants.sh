set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention.py -x -q -m gpu > gpurun_out/r40_pytest_attn.txt 2>&1; tail -5 gpurun_out/r40_pytest_attn.txt
for v in "" _poly25 _poly37 _poly50 _hilo; do
  echo "== variant '$v'" >> gpurun_out/r40_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r40_attn_case.txt 2>&1
done
cat gpurun_out/r40_attn_case.txt
