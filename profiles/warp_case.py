"""Launches ff_warp_affine_blend on the L2-exceeding roofline batch (N*C = 32768 channels of 64x64 fp32) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops
dev = torch.device("cuda:0")
NC = 32768
src = torch.randn(1, NC, 64, 64, device=dev)
bg = torch.randn(1, NC, 64, 64, device=dev)
mask = (torch.rand(1, 64, 64, device=dev) > 0.5).to(torch.uint8)
th = torch.tensor([[[0.95, 0.2, 0.05], [-0.2, 0.95, -0.03]]], device=dev)
out = torch.empty_like(bg)
for _ in range(3):
    ops.warp_affine_blend(src, th, mask_src=mask, bg=bg, out=out)
torch.cuda.synchronize()
