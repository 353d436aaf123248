"""Diagnostic (not a test): for the fuzz seeds of test_random_plans_vs_plan_interpreter print the worst (stream, head)
and its plan entry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from freefine_b200 import ops, plans
from oracle import ff_oracle as O
from plan_interp import run_plan
dev = torch.device("cuda:0")
for seed in range(24):
    rng = np.random.default_rng(9000 + seed)
    B = int(rng.integers(1, 4)); heads = int(rng.choice([1, 2, 4])); d = int(rng.choice([8, 16, 24, 40, 64, 80, 96, 160]))
    Sq, Skv = int(rng.integers(1, 261)), int(rng.integers(1, 321))
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Sq, heads * d, generator=g).bfloat16().float()
    k = torch.randn(B, Skv, heads * d, generator=g).bfloat16().float()
    v = (torch.randn(B, Skv, heads * d, generator=g) * 2).bfloat16().float()
    S = max(Sq, Skv)
    dens = [0.0, 1.0, float(rng.uniform(0.05, 0.95)), float(rng.uniform(0.3, 0.7))]
    rng.shuffle(dens)
    flat = [(rng.random(S) < p_).astype(np.uint8) for p_ in dens]
    plan = plans._empty(B, heads)
    for s_ in range(B):
        for h in range(heads):
            for _ in range(int(rng.integers(1, 4))):
                kv = int(rng.integers(0, B)); km = int(rng.integers(-1, 4)); rm = int(rng.integers(-1, 4)); flags = 0
                if km >= 0 and rng.random() < 0.4: flags |= plans.FF_PASS_KEY_INVERT
                if km >= 0 and rm >= 0 and rng.random() < 0.5: flags |= plans.FF_PASS_ROW_XOR
                if rm >= 0 and rng.random() < 0.3: flags |= plans.FF_PASS_ROW_WEIGHT
                kv2 = -1
                if not (flags & plans.FF_PASS_ROW_XOR) and rng.random() < 0.25: kv2 = int(rng.integers(0, B))
                plans._add(plan, s_, h, kv, float(rng.uniform(0.1, 1.0)), key_mask=km, row_mask=rm, flags=flags, kv2=kv2, key_mask2=-1)
    words = ops.mask_words(S)
    arr = np.zeros((len(flat), words), np.uint32)
    for i, m in enumerate(flat):
        b = O.pack_bits(m != 0); arr[i, :len(b)] = b
    ref = run_plan(q, k, v, plan, heads, d ** -0.5, arr)
    bm = torch.from_numpy(arr.view(np.int32)).to(dev)
    pc = torch.tensor([int(m[:Skv].sum()) for m in flat], dtype=torch.int32, device=dev)
    out = ops.attn_masked_kv(q.to(dev).bfloat16(), k.to(dev).bfloat16(), v.to(dev).bfloat16(), ops.to_device_bytes(plan, dev), heads,
                             d ** -0.5, bm, pc, out_dtype=torch.float32, p_operand="bf16x2")
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().reshape(B, Sq, heads, d).amax((1, 3))
    print(f"seed {seed}: B={B} H={heads} d={d} Sq={Sq} Skv={Skv} dens={['%.2f' % x for x in dens]} pop={[int(m[:Skv].sum()) for m in flat]} max err {float(err.max()):.3e}")
    if float(err.max()) > 6e-3:
        for s_ in range(B):
            for h in range(heads):
                if float(err[s_, h]) > 6e-3:
                    e = plan[s_, h]
                    desc = [(int(ps['kv_stream']), int(ps['key_mask']), int(ps['row_mask']), int(ps['flags']), round(float(ps['weight']), 2), int(ps['kv_stream2'])) for ps in e['passes'][:int(e['n_pass'])]]
                    print(f"   (s={s_},h={h}) err {float(err[s_, h]):.3f}  passes (kv,kmask,rmask,flags,w,kv2): {desc}")
