"""TEST INFRASTRUCTURE -- a literal CPU interpreter (torch fp32) of the FFAttnHeadPlan semantics documented in
include/freefine_b200.h.  It lets the CPU suite check the host-side plan builders (freefine_b200/plans.py) against
the oracle without a GPU, and gives the GPU tests a second, slow, reading of what ff_attn_masked_kv must compute.
Never imported by the product."""
import numpy as np
import torch

from freefine_b200.plans import (FF_PASS_KEY2_INVERT, FF_PASS_KEY2_PREFIX, FF_PASS_KEY_INVERT, FF_PASS_KEY_PREFIX,
                                 FF_PASS_ROW_WEIGHT, FF_PASS_ROW_XOR)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    w = np.asarray(words).astype(np.uint32)
    return ((w[:, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(-1)[:n].astype(bool)


def run_plan(q, k, v, plan, heads, scale, bitmasks=None):
    """q [B,Sq,C], k/v [Bk,Skv,C] fp32; plan: numpy PLAN_DTYPE [B,heads]; bitmasks: numpy uint32 [n, words]."""
    B, Sq, C = q.shape
    Skv = k.shape[1]
    d = C // heads
    out = torch.zeros(B, Sq, C)
    qh = q.reshape(B, Sq, heads, d)
    kh = k.reshape(k.shape[0], Skv, heads, d)
    vh = v.reshape(v.shape[0], Skv, heads, d)

    def kbits(mid, S, prefix=False):
        if mid < 0:
            return torch.ones(S, dtype=torch.bool)
        b = torch.from_numpy(unpack_bits(bitmasks[mid], S))
        return (torch.arange(S) < int(b.sum())) if prefix else b      # PREFIX: keys were sorted "set bits first"

    for s in range(B):
        for h in range(heads):
            e = plan[s, h]
            acc = torch.zeros(Sq, d)
            for ps in e["passes"][: int(e["n_pass"])]:
                flags = int(ps["flags"])
                rb = (torch.zeros(Sq, dtype=torch.bool) if ps["row_mask"] < 0
                      else torch.from_numpy(unpack_bits(bitmasks[int(ps["row_mask"])], Sq)))
                segs = [(int(ps["kv_stream"]), int(ps["key_mask"]), bool(flags & FF_PASS_KEY_INVERT), bool(flags & FF_PASS_KEY_PREFIX))]
                if ps["kv_stream2"] >= 0:
                    segs.append((int(ps["kv_stream2"]), int(ps["key_mask2"]), bool(flags & FF_PASS_KEY2_INVERT),
                                 bool(flags & FF_PASS_KEY2_PREFIX)))
                ks, vs, al = [], [], []
                for kv, km, inv, pfx in segs:
                    ks.append(kh[kv, :, h])
                    vs.append(vh[kv, :, h])
                    a = kbits(km, Skv, pfx)[None, :].expand(Sq, Skv)
                    if km >= 0:                      # a segment without a key mask admits every key, whatever the flags
                        if inv:
                            a = ~a
                        if flags & FF_PASS_ROW_XOR:
                            a = a ^ rb[:, None]
                    al.append(a)
                kk, vv, allowed = torch.cat(ks), torch.cat(vs), torch.cat(al, 1)
                sc = (qh[s, :, h] @ kk.T) * scale
                empty = ~allowed.any(1, keepdim=True)
                sc = torch.where(allowed | empty, sc, torch.full_like(sc, float("-inf")))
                sc = torch.where(empty.expand_as(sc), torch.zeros_like(sc), sc)       # quirk Q4: uniform
                o = torch.softmax(sc, -1) @ vv
                w = float(ps["weight"])
                if flags & FF_PASS_ROW_WEIGHT:
                    o = o * rb[:, None].float()
                acc = acc + w * o
            out[s, :, h * d:(h + 1) * d] = acc
    return out
