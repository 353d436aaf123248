"""Per-tile clock64 timeline of two CTAs of the ring attention kernel (library built with -DFF_TIMELINE):
    python -m freefine_b200.csrc.build --variant=tl -DFF_TIMELINE
    FREEFINE_B200_LIB=freefine_b200/lib/libfreefine_b200_tl.so FF_ATTN_RING=1 python profiles/timeline_ring.py [S d]
Roles: 0/1 = warp 0 of softmax warpgroup 0/1, 2 = QK^T issuer, 3 = P.V issuer (csrc/attn_ring.cuh, RT_TL sites).
softmax sites: 0 wait s_full, 1 s_full seen, 2 row loaded, 3 max done, 4 exp done, 5 P stored + arrived
issuer sites : 0 loop top, 1 k_full/v_full seen, 2 pv_done/p_full seen, 3 MMAs + commits issued"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans, synth, _lib

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 40
E, heads = 8, 8
hw = int(S ** 0.5)
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
buf = torch.zeros(2 * 4 * 64 * 8, dtype=torch.int64, device=dev)
_lib.check(_lib.load().ff_debug_set_timeline(buf.data_ptr()), "timeline")
ops.attn_masked_kv(q, k, v, plan, heads, d ** -0.5, bits, pop, out=out)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(2, 4, 64, 8)
names = ["softmax wg0", "softmax wg1", "QK issuer", "PV issuer"]
for cta in range(2):
    t0 = int(min(x for x in t[cta].reshape(-1) if x > 0)) if (t[cta] > 0).any() else 0
    print(f"== CTA {cta} (cycles relative to the CTA's first stamp)")
    for role in range(4):
        r = t[cta, role]
        tiles = [i for i in range(8, 56) if r[i, 0] > 0]
        if len(tiles) < 4:
            continue
        nsite = 6 if role < 2 else 4
        print(f"-- {names[role]}: tile " + " ".join(f"site{s}" for s in range(nsite)))
        for i in tiles[:14]:
            print(f"   {i:3d} " + " ".join(f"{int(r[i, s]) - t0:7d}" for s in range(nsite)))
        per = np.diff(np.array([r[i, 0] for i in tiles], dtype=np.int64))
        seg = np.array([[int(r[i, s + 1]) - int(r[i, s]) for s in range(nsite - 1)] for i in tiles])
        print(f"   tile period (this role): mean {per.mean():.0f} min {per.min()} max {per.max()}  | mean phase durations: "
              + " ".join(f"{x:.0f}" for x in seg.mean(0)))
    # cross-role latencies for tiles seen by softmax wg0
    sm, qk, pv = t[cta, 0], t[cta, 2], t[cta, 3]
    lat = [(int(pv[i, 2]) - int(sm[i, 5]), int(pv[i, 3]) - int(pv[i, 2])) for i in range(8, 56) if sm[i, 5] > 0 and pv[i, 2] > 0]
    if lat:
        a = np.array(lat)
        print(f"   softmax arrive -> PV issuer sees p_full: mean {a[:, 0].mean():.0f}; PV issue: mean {a[:, 1].mean():.0f}")
