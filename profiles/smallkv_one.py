"""One ff_attn_plain_smallkv shape, a few plain launches -- target of an ncu capture.  usage: smallkv_one.py B S_q S_kv d"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops
B, s_q, s_kv, d = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda:0")
q = torch.randn(B, s_q, 8 * d, device=dev).bfloat16()
k = torch.randn(B, s_kv, 8 * d, device=dev).bfloat16()
v = torch.randn(B, s_kv, 8 * d, device=dev).bfloat16()
for _ in range(3):
    o = ops.attn_plain_smallkv(q, k, v, 8, d ** -0.5)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
