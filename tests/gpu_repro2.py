import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans
from oracle import cases, ff_oracle as O
dev = torch.device("cuda:0")
S, d, res, method, kind, prefix = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6] == "1"
heads, E = 8, 2
q, k, v = cases.qkv(4 * E, S, heads * d, 500 + S + d)
flat = []
for e in range(E):
    flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 700 + e)), S).numpy())
    flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 800 + e)), S).numpy())
cg = 0.55
plan = plans.tca_plan(E, heads, method, cg, lambda e: 2 * e, lambda e: 2 * e + 1, kind=kind, prefix=prefix)
words = ops.mask_words(S)
arr = np.zeros((len(flat), words), np.uint32)
for i, m in enumerate(flat):
    b = O.pack_bits(np.asarray(m) != 0); arr[i, :len(b)] = b
bm = torch.from_numpy(arr.view(np.int32)).to(dev)
pc = torch.tensor([int((np.asarray(m) != 0).sum()) for m in flat], dtype=torch.int32, device=dev)
kk, vv = k, v
if prefix:
    idx = plans.kv_sort_index(torch.from_numpy(np.stack([np.asarray(m) != 0 for m in flat])), [2 * (s // 4) if s % 2 else -1 for s in range(4 * E)])
    kk = k.reshape(-1, k.shape[-1])[idx].reshape(k.shape); vv = v.reshape(-1, v.shape[-1])[idx].reshape(v.shape)
for rep in range(3):
    out = ops.attn_masked_kv(q.to(dev).bfloat16(), kk.to(dev).bfloat16(), vv.to(dev).bfloat16(), ops.to_device_bytes(plan, dev),
                             heads, d ** -0.5, bm, pc, out_dtype=torch.float32)
    torch.cuda.synchronize()
    out = out.cpu()
    errs = []
    for e in range(E):
        ref = O.tca(q[4*e:4*e+4], k[4*e:4*e+4], v[4*e:4*e+4], heads, d ** -0.5, flat[2*e], flat[2*e+1], method, cg, kind=kind)
        errs.append((out[4*e:4*e+4] - ref).abs().reshape(4, S, heads, d))
    err = torch.cat(errs)
    print(f"rep {rep}: max err {float(err.max()):.3e}")
    if float(err.max()) > 2e-3:
        print(" per (stream, head):\n", np.array2string(err.amax((1, 3)).numpy(), precision=2, suppress_small=True))
        s_, h_ = np.unravel_index(int(err.amax((1, 3)).argmax()), (4 * E, heads))
        rows = err[s_, :, h_].amax(1)
        print(f" worst (s={s_},h={h_}) per q-tile(128) max:", np.array2string(rows.reshape(-1, 128).amax(1).numpy(), precision=2))
        bad = (rows > 2e-3).nonzero().flatten()
        print("  #bad rows", len(bad), "first", bad[:10].tolist(), " tgt bits of those", [int(flat[2*(s_//4)+1][r]) for r in bad[:10].tolist()])
print("masks popcounts", [int(m.sum()) for m in flat])
