"""One GroupNorm+SiLU shape, a handful of plain launches -- the target of an ncu capture (profiles/scripts/round2c_gn_ncu.sh).
usage: gn_one.py N C H W [launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops

n, c, h, w = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
x = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
ga, be = torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16()
add = torch.randn(n, c, device=dev)
for _ in range(reps):
    y = ops.group_norm_nhwc(x, ga, be, 32, 1e-5, add_nc=add, silu=True)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
