#!/bin/bash
# ncu --set full capture of the dominant attention launch (profiles/attn_case.py: S=4096, d=40, 32 streams) for one
# FF_ATTN_RING mode; the .ncu-rep comes back in gpurun_out/ and is read here with ncu -i ... --page raw/source --csv
mkdir -p gpurun_out
M=${1:-3}; T=${2:-e}
FF_ATTN_RING=$M timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 1 -c 1 \
   -o gpurun_out/${T}_attn_ring${M} -f python profiles/attn_case.py 1 > gpurun_out/${T}_ncu_attn_ring${M}.log 2>&1
tail -3 gpurun_out/${T}_ncu_attn_ring${M}.log
