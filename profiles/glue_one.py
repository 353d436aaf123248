"""GEGLU and LayerNorm at the 64 x 64 sizes of a 32-stream call, a few plain launches each -- target of an ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops
dev = torch.device("cuda:0")
h = torch.randn(32, 4096, 2560, device=dev).bfloat16()
x = torch.randn(32, 4096, 320, device=dev).bfloat16()
ga, be = torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16()
for _ in range(3):
    o1 = ops.geglu(h)
    o2 = ops.layer_norm(x, ga, be, 1e-5)
torch.cuda.synchronize()
print("ok", float(o1.float().abs().mean()), float(o2.float().abs().mean()))
