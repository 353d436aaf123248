"""Per-tile clock64 timeline of two CTAs of the S=4096 d=40 attention launch (library built with -DFF_TIMELINE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans, synth, _lib

dev = torch.device("cuda:0")
E, heads, S, d = 8, 8, 4096, 40
hw = 64
g = torch.Generator().manual_seed(0)
q = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
k = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
v = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
masks = []
for e in range(E):
    b = synth.make_edit(e, 512)
    masks += [b["mask"], np.roll(b["mask"], (20, -30), (0, 1))]
bits, pop = ops.mask_downsample_pack(torch.from_numpy(np.stack(masks)).to(dev), hw, hw)
plan_np = plans.tca_plan(E, heads, "tca", 0.5, lambda e: 2 * e, lambda e: 2 * e + 1, prefix=True)
plan = ops.to_device_bytes(plan_np, dev)
shifts = torch.arange(32, device=dev, dtype=torch.int32)
key_bits = ((bits[:, :, None] >> shifts) & 1).reshape(bits.shape[0], -1)[:, :S]
idx = plans.kv_sort_index(key_bits, [2 * (s // 4) if s % 2 else -1 for s in range(4 * E)])
k, v = ops.kv_gather_cast(k, v, heads, idx)
out = ops.attn_masked_kv(q, k, v, plan, heads, d ** -0.5, bits, pop)
torch.cuda.synchronize()
buf = torch.zeros(2 * 2 * 64 * 8, dtype=torch.int64, device=dev)
_lib.check(_lib.load().ff_debug_set_timeline(buf.data_ptr()), "timeline")
ops.attn_masked_kv(q, k, v, plan, heads, d ** -0.5, bits, pop, out=out)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(2, 2, 64, 8)
for cta in range(2):
    sm, mm = t[cta, 0], t[cta, 1]
    t0 = sm[0, 0]
    print(f"== CTA {cta}: softmax warp 0 (cycles rel. to its first stamp): wait_begin s_full loaded max_done exp_done st_waited arrived | MMA: kvwait_begin kv_full qk_issued p_full pv_issued")
    for i in range(4, 40):
        a = [int(x - t0) if x else -1 for x in sm[i, :7]]
        b = [int(x - t0) if x else -1 for x in mm[i, :5]]
        print(f"tile {i:2d} S " + " ".join(f"{x:7d}" for x in a) + "  | M " + " ".join(f"{x:7d}" for x in b))
    rows = [i for i in range(4, 60) if sm[i, 0] and sm[i, 6]]              # this warp's tiles (one parity)
    dt = np.diff(sm[rows, 0].astype(np.int64))
    print("softmax tile period (own tiles): mean %.0f  min %d  max %d" % (dt.mean(), dt.min(), dt.max()))
    seg = (sm[rows, 1:7] - sm[rows, 0:6]).astype(np.int64)
    print("mean phase durations: wait_s %.0f  load+max(h0) %.0f  load+max(h1) %.0f  exp %.0f  st_wait %.0f  arrive %.0f" % tuple(seg.mean(0)))
    for i in range(8, 20):
        a = [int(x - t0) if x else -1 for x in sm[i, :7]]
        b = [int(x - t0) if x else -1 for x in mm[i, :5]]
        print(f"tile {i:2d} S " + " ".join(f"{x:7d}" for x in a) + "  | M " + " ".join(f"{x:7d}" for x in b))
    mrows = [i for i in range(4, 60) if mm[i, 0] and mm[i, 4]]
    if mrows:
        md = np.diff(mm[mrows, 0].astype(np.int64))
        print("issuer tile period: mean %.0f; p_full wait %.0f, issue %.0f" % (md.mean(), (mm[mrows, 3] - mm[mrows, 0]).mean(), (mm[mrows, 4] - mm[mrows, 3]).mean()))
        # hand-shake: softmax arrive(t) -> issuer saw p_full(t); issuer issued -> softmax saw s_full(t+2)
        hs1 = [int(mm[i, 3] - sm[i, 6]) for i in rows if mm[i, 3]]
        hs2 = [int(sm[i + 2, 1] - mm[i, 4]) for i in rows if i + 2 < 64 and sm[i + 2, 1] and mm[i, 4]]
        print("arrive(t) -> issuer resumed: mean %.0f | PV(t),QK(t+2) issued -> softmax resumed on s_full(t+2): mean %.0f" % (np.mean(hs1), np.mean(hs2)))
