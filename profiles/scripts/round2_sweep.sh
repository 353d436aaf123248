#!/bin/bash
# Config 3 on hardware: a sweep of M independent edits sharded i mod W with the final latent gather over NCCL inside the
# timed region (bench.py --sweep), non-divisible M so that the sampler's tail padding + de-duplication is exercised.
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/scripts/round2_sweep.sh 2 61'
N=${1:-2}; M=${2:-61}
mkdir -p gpurun_out
python bench.py --sweep $M --start-step 35 --warmup 3 > gpurun_out/s_sweep_n1_m$M.json 2> gpurun_out/s_sweep_n1.err; tail -2 gpurun_out/s_sweep_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N \
   --sweep $M --start-step 35 --warmup 3 > gpurun_out/s_sweep_n${N}_m$M.json 2> gpurun_out/s_sweep_n$N.err; tail -3 gpurun_out/s_sweep_n$N.err
python - <<PY
import json
for f in ("gpurun_out/s_sweep_n1_m$M.json", "gpurun_out/s_sweep_n${N}_m$M.json"):
    try:
        d = json.load(open(f)); print(f, d["n_gpus"], round(d["value"], 3), "edits/s", d["sweep"])
    except Exception as e:
        print(f, "ERR", e)
PY
