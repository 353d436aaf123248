"""Diagnostic script (not a test): runs tiny attention cases on the GPU and prints per-(stream, head) error maps so
that a layout / descriptor mistake in the tcgen05 kernel can be localised from one gpurun call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans
from oracle import ff_oracle as O

dev = torch.device("cuda:0")
torch.manual_seed(0)
P_OP = os.environ.get("FF_P", "f16")
print("P operand:", P_OP)


def _v(v, heads):
    v = v.to(dev).bfloat16().contiguous()
    return ops.kv_gather_cast(v, v, heads, None, p_operand=P_OP)[1]


def run(name, B, S, Skv, heads, d, q=None, k=None, v=None):
    C = heads * d
    g = torch.Generator().manual_seed(1)
    q = (torch.randn(B, S, C, generator=g)).bfloat16().float() if q is None else q
    k = (torch.randn(B, Skv, C, generator=g)).bfloat16().float() if k is None else k
    v = (torch.randn(B, Skv, C, generator=g)).bfloat16().float() if v is None else v
    pl = ops.to_device_bytes(plans.plain_plan(B, heads), dev)
    try:
        out = ops.attn_masked_kv(q.to(dev).bfloat16(), k.to(dev).bfloat16(), _v(v, heads), pl, heads, d ** -0.5,
                                 out_dtype=torch.float32)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{name}] FAILED: {e}")
        return None
    out = out.cpu()
    ref = O.plain_attention(q, k, v, heads, d ** -0.5)
    err = (out - ref).abs()
    print(f"[{name}] B={B} S={S} Skv={Skv} H={heads} d={d}: max err {float(err.max()):.3e}  nan={int(torch.isnan(out).sum())}")
    if float(err.max()) > 2e-3 or torch.isnan(out).any():
        eh = err.reshape(B, S, heads, d)
        print("  per (stream,head) max err:\n", np.array2string(eh.amax((1, 3)).numpy(), precision=3))
        print("  per row-block(32) max err (stream0,head0):", np.array2string(eh[0, :, 0].amax(1).reshape(-1, 32).amax(1).numpy(), precision=3))
        print("  per column max err (stream0,head0):", np.array2string(eh[0, :, 0].amax(0).numpy(), precision=3))
        print("  out[0,0,:8]", out[0, 0, :8].numpy(), "\n  ref[0,0,:8]", ref[0, 0, :8].numpy())
    return out


# 1) V = identity-like probes: uniform attention (q=0) returns the column mean of V
S = 128
for d in (16, 40, 64, 80, 160):
    heads = 2
    C = heads * d
    q = torch.zeros(1, S, C)
    k = torch.zeros(1, S, C)
    v = torch.arange(S, dtype=torch.float32)[None, :, None].expand(1, S, C).contiguous() / 64
    v = v + torch.arange(C, dtype=torch.float32)[None, None, :] / 8
    run(f"uniform d={d}", 1, S, S, heads, d, q, k, v.bfloat16().float())
# 2) random, single tile / multi tile / ragged
run("rand 1 tile d16", 1, 128, 128, 2, 16)
run("rand 1 tile d40", 1, 128, 128, 2, 40)
run("rand 2 tiles d40", 2, 256, 256, 2, 40)
run("rand 4 tiles d80", 1, 512, 512, 2, 80)
run("rand ragged d40", 1, 200, 77, 2, 40)
run("rand d160", 1, 256, 256, 2, 160)
run("rand big d40", 4, 1024, 1024, 8, 40)

# 3) masked two-pass plans (TCA) against the oracle, per (stream, head) error map
from oracle import cases


def run_tca(name, heads, d, S, res, method="tca", kind="edit", cg=0.6):
    q, k, v = cases.qkv(4, S, heads * d, 11)
    src = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 211)), S).numpy()
    tgt = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 111)), S).numpy()
    words = ops.mask_words(S)
    arr = np.zeros((2, words), np.uint32)
    for i, m in enumerate((src, tgt)):
        b = O.pack_bits(m != 0)
        arr[i, :len(b)] = b
    bm = torch.from_numpy(arr.view(np.int32)).to(dev)
    pc = torch.tensor([int((src != 0).sum()), int((tgt != 0).sum())], dtype=torch.int32, device=dev)
    plan = plans.tca_plan(1, heads, method, cg, lambda e: 0, lambda e: 1, kind=kind)
    try:
        out = ops.attn_masked_kv(q.to(dev).bfloat16(), k.to(dev).bfloat16(), _v(v, heads),
                                 ops.to_device_bytes(plan, dev), heads, d ** -0.5, bm, pc, out_dtype=torch.float32)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{name}] FAILED: {e}")
        return
    ref = O.tca(q, k, v, heads, d ** -0.5, src, tgt, method, cg, kind=kind)
    err = (out.cpu() - ref).abs().reshape(4, S, heads, d)
    print(f"[{name}] H={heads} d={d} S={S} {method}/{kind}: max err {float(err.max()):.3e} nan={int(torch.isnan(out).sum())}")
    if float(err.max()) > 2e-3:
        print("  per (stream,head) max err:\n", np.array2string(err.amax((1, 3)).numpy(), precision=3))


run_tca("tca d8", 8, 8, 256, 128)
run_tca("tca d40", 8, 40, 256, 128)
run_tca("mmsa d40", 8, 40, 256, 128, method="mmsa")
run_tca("bg d80", 8, 80, 256, 128, kind="bg")
run_tca("tca d160", 2, 160, 64, 64)
