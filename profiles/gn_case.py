"""GroupNorm+SiLU kernel pair at the sizes of a 32- / 16-stream UNet call, timed as CUDA-graph replays (launch cost
excluded); FF_GN_CHUNK_PX selects the pixels per CTA (tuning knob of csrc/unet_glue.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops

dev = torch.device("cuda:0")
for (n, c, h, w) in ((32, 320, 64, 64), (16, 320, 64, 64), (32, 960, 64, 64), (32, 640, 32, 32), (32, 1280, 16, 16), (16, 320, 96, 96)):
    x = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    ga, be = torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16()
    add = torch.randn(n, c, device=dev)
    fn = lambda: ops.group_norm_nhwc(x, ga, be, 32, 1e-5, add_nc=add, silu=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    byt = 2 * x.numel() * 2
    print(f"chunk_px={os.environ.get('FF_GN_CHUNK_PX', 'default')} {n}x{c}x{h}x{w}: {best * 1e3:7.1f} us  {byt / best / 1e6:6.0f} GB/s algorithmic")
