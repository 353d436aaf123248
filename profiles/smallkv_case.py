"""Plain attention over short key sequences at the shapes of a 32- / 16-stream UNet call (text cross-attention: 77 keys;
8 x 8 self-attention: 64 keys): ff_attn_plain_smallkv against the same layers through ff_kv_gather_cast + ff_attn_masked_kv
(the path they took before round 2c).  CUDA-graph replays of 10 launches, us per launch; algorithmic bytes = Q in + O out."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from freefine_b200 import ops, plans

dev = torch.device("cuda:0")


def timed(fn):
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
    return best * 1e3


for (B, s_q, s_kv, d) in ((32, 4096, 77, 40), (16, 4096, 77, 40), (32, 1024, 77, 80), (16, 1024, 77, 80), (32, 256, 77, 160),
                          (32, 64, 77, 160), (16, 64, 64, 160), (16, 9216, 77, 40), (32, 256, 256, 160), (16, 256, 256, 160)):
    heads = 8
    C = heads * d
    q = torch.randn(B, s_q, C, device=dev).bfloat16()
    k = torch.randn(B, s_kv, C, device=dev).bfloat16()
    v = torch.randn(B, s_kv, C, device=dev).bfloat16()
    sc = d ** -0.5
    plan = ops.to_device_bytes(plans.plain_plan(B, heads), dev)

    def old():
        kk, vv = ops.kv_gather_cast(k, v, heads, None, p_operand="f16")
        return ops.attn_masked_kv(q, kk, vv, plan, heads, sc, None, None, out_dtype=torch.bfloat16)

    new = lambda: ops.attn_plain_smallkv(q, k, v, heads, sc, out_dtype=torch.bfloat16)
    err = float((old().float() - new().float()).abs().max())
    t_old, t_new = timed(old), timed(new)
    byt = 2 * q.numel() * 2
    print(f"streams={B} S_q={s_q} S_kv={s_kv} d={d}: masked_kv path {t_old:7.1f} us   smallkv {t_new:7.1f} us   "
          f"({byt / t_new / 1e3:6.0f} GB/s algorithmic)   max |diff| {err:.2e}")
