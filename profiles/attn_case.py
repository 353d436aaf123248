"""Launches the bench's dominant kernel in isolation (for `ncu --set full`): ff_attn_masked_kv on the up3 TCA layer
shape of 8 batched 512x512 edits -- 32 streams x 8 heads, S=4096, d=40, 'tca' plans with synthetic GeoBench-like
masks -- and the up2 shape (S=1024, d=80).  Prints CUDA-event timings (never quote numbers taken under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans, synth

dev = torch.device("cuda:0")
E, heads = 8, 8
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for S, d in ((4096, 40), (1024, 80)):
    hw = int(S ** 0.5)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    k = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    v = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    masks = []
    for e in range(E):
        b = synth.make_edit(e, 512)
        masks += [b["mask"], np.roll(b["mask"], (20, -30), (0, 1))]          # src (fg_ref), tgt (fg_retain)
    bits, pop = ops.mask_downsample_pack(torch.from_numpy(np.stack(masks)).to(dev), hw, hw)
    prefix = os.environ.get("FF_PREFIX", "1") == "1"       # the controller's default: keys sorted "source first"
    plan_np = plans.tca_plan(E, heads, "tca", 0.5, lambda e: 2 * e, lambda e: 2 * e + 1, prefix=prefix)
    plain = os.environ.get("FF_PLAIN", "0") == "1"         # the plain self-attention layers (outside the TCA scope / inversion)
    if plain:
        plan_np, prefix = plans.plain_plan(4 * E, heads), False
    plan = ops.to_device_bytes(plan_np, dev)
    idx = None
    if prefix:
        shifts = torch.arange(32, device=dev, dtype=torch.int32)
        key_bits = ((bits[:, :, None] >> shifts) & 1).reshape(bits.shape[0], -1)[:, :S]
        idx = plans.kv_sort_index(key_bits, [2 * (s // 4) if s % 2 else -1 for s in range(4 * E)])
    k, v = ops.kv_gather_cast(k, v, heads, idx, p_operand=os.environ.get("FF_P", "f16"))
    flops = plans.algorithmic_flops(plan_np, S, S, d, pop.cpu().numpy())
    out = ops.attn_masked_kv(q, k, v, plan, heads, d ** -0.5, bits, pop)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        ops.attn_masked_kv(q, k, v, plan, heads, d ** -0.5, bits, pop, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    print(f"S={S} d={d} streams={4*E} {'PLAIN ' if plain else ''}prefix={prefix} P={os.environ.get('FF_P', 'f16')}: {min(ms):.3f} ms best / {sum(ms)/len(ms):.3f} ms avg, algorithmic {flops/1e9:.1f} GFLOP -> "
          f"{flops / (min(ms) * 1e-3) / 1e12:.1f} TFLOP/s (dense-equivalent {7*4*S*S*d*heads*E/1e9:.1f} GFLOP)")
