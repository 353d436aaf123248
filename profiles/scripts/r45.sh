mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_kernels.py -q -m gpu > gpurun_out/r45_pytest.txt 2>&1; tail -4 gpurun_out/r45_pytest.txt
for v in "" _komufu _komax _koboth _spin _wg2 _poly25; do
  echo "== variant '$v'" >> gpurun_out/r45_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r45_attn_case.txt 2>&1
done
FF_P=bf16x2 timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r45_attn_case.txt 2>&1
grep -v "^S=1024" gpurun_out/r45_attn_case.txt
