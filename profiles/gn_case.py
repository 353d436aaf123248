"""GroupNorm+SiLU kernel pair at the sizes of a 32- / 16-stream UNet call, timed as CUDA-graph replays (launch cost
excluded); FF_GN_CHUNK_PX selects the pixels per CTA (tuning knob of csrc/unet_glue.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops

dev = torch.device("cuda:0")
SHAPES = ((32, 320, 64, 64), (16, 320, 64, 64), (32, 960, 64, 64), (32, 640, 32, 32), (32, 1280, 16, 16), (16, 320, 96, 96),
          # the small levels of a 32- / 16-stream call (register-resident kernel unless FF_GN_SMALL=0)
          (32, 1280, 8, 8), (32, 2560, 8, 8), (16, 1280, 8, 8), (32, 2560, 16, 16), (32, 1920, 16, 16), (16, 1280, 16, 16),
          (32, 1280, 32, 32), (32, 320, 32, 32), (32, 1920, 32, 32), (32, 960, 32, 32))
for (n, c, h, w) in SHAPES:
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
    print(f"chunk_px={os.environ.get('FF_GN_CHUNK_PX', 'default')} small={os.environ.get('FF_GN_SMALL', '1')} {n}x{c}x{h}x{w}: {best * 1e3:7.1f} us  {byt / best / 1e6:6.0f} GB/s algorithmic")

# LayerNorm at the transformer widths of a 32-stream call (sub-warp-row kernel for C = 320 / 640 / 1280)
for (m, c) in ((32 * 4096, 320), (32 * 1024, 640), (32 * 256, 1280), (16 * 4096, 320)):
    x = torch.randn(1, m, c, device=dev).bfloat16()
    ga, be = torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16()
    fn = lambda: ops.layer_norm(x, ga, be, 1e-5)
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
    print(f"layer_norm {m}x{c}: {best * 1e3:7.1f} us  {byt / best / 1e6:6.0f} GB/s algorithmic")

# GEGLU at the feed-forward widths of a 32-stream call: h [M, 2F] -> [M, F]
for (m, f) in ((32 * 4096, 1280), (32 * 1024, 2560), (32 * 256, 5120), (16 * 4096, 1280)):
    h = torch.randn(1, m, 2 * f, device=dev).bfloat16()
    fn = lambda: ops.geglu(h)
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
    byt = 3 * m * f * 2
    print(f"geglu {m}x{f}: {best * 1e3:7.1f} us  {byt / best / 1e6:6.0f} GB/s algorithmic")
