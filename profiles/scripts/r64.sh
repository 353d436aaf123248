mkdir -p gpurun_out
for v in "" _nohint _s2 _s2sl _s3sl; do
  echo "== variant '$v'" >> gpurun_out/r64_attn_case.txt
  FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200$v.so timeout 120 python profiles/attn_case.py 5 >> gpurun_out/r64_attn_case.txt 2>&1
done
cat gpurun_out/r64_attn_case.txt
