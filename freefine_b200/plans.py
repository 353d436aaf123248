"""Host-side construction of the per-(stream, head) attention plans consumed by ff_attn_masked_kv.

This is where the reference's mask-building *semantics* live (pure host logic, no arithmetic on activations):

* stream layout `[u_e, u_r, c_e, c_r]` per edit and the KV replacement `[u_r, u_r, c_r, c_r]`
  (cross_manner_attention_modulate, src/utils/attention.py:1033-1035);
* quirk Q0: the 4-entry mask stack `[M, 1, M, 1]` is *tiled* over batch*heads by `.repeat(heads,1,1)`
  (attention.py:859,881) while q/k/v are batch-major (head_to_batch_dim :758-767), so batch*head index
  j = heads*s + h receives stack entry j mod 4: only entries 0 and 2 carry the region masks;
* which rows read which keys (tgt rows -> source-object keys, other rows -> the complement, :1069/:1081),
  the bg-gen variant (:1284-1324), the N-source compose variant (:1092-1140) and SSA/SDSA (:1142-1192).

Mask ids index rows of the bit-vector table the caller passes along (see controller: one table per token-grid
resolution, same ids at every resolution, so one plan serves every layer of a step).
"""
from __future__ import annotations

import numpy as np

from ._lib import (FF_MAX_PASS, FF_PASS_KEY2_INVERT, FF_PASS_KEY2_PREFIX, FF_PASS_KEY_INVERT, FF_PASS_KEY_PREFIX,
                   FF_PASS_ROW_WEIGHT, FF_PASS_ROW_XOR, PLAN_BYTES)

PASS_DTYPE = np.dtype([("kv_stream", "<i4"), ("key_mask", "<i4"), ("row_mask", "<i4"), ("flags", "<u4"),
                       ("weight", "<f4"), ("kv_stream2", "<i4"), ("key_mask2", "<i4"), ("reserved", "<i4")])
PLAN_DTYPE = np.dtype([("n_pass", "<i4"), ("reserved", "<i4", (3,)), ("passes", PASS_DTYPE, (FF_MAX_PASS,))])
assert PLAN_DTYPE.itemsize == PLAN_BYTES


def q0_masked(heads: int, stream_in_edit: int, head: int) -> bool:
    """Quirk Q0: does (stream, head) receive the region masks?  (attention.py:859,881 vs :758-767)"""
    return ((heads * stream_in_edit + head) % 4) in (0, 2)


def _empty(n_streams: int, heads: int) -> np.ndarray:
    p = np.zeros((n_streams, heads), PLAN_DTYPE)
    p["passes"]["kv_stream"] = -1
    p["passes"]["kv_stream2"] = -1
    p["passes"]["key_mask"] = -1
    p["passes"]["key_mask2"] = -1
    p["passes"]["row_mask"] = -1
    return p


def _add(plan, s, h, kv, weight, key_mask=-1, row_mask=-1, flags=0, kv2=-1, key_mask2=-1):
    e = plan[s, h]
    i = int(e["n_pass"])
    if i >= FF_MAX_PASS:
        raise ValueError(f"more than {FF_MAX_PASS} passes for stream {s} head {h}")
    ps = e["passes"][i]
    ps["kv_stream"], ps["key_mask"], ps["row_mask"], ps["flags"] = kv, key_mask, row_mask, flags
    ps["weight"], ps["kv_stream2"], ps["key_mask2"] = weight, kv2, key_mask2
    plan[s, h]["n_pass"] = i + 1


def plain_plan(n_streams: int, heads: int, kv_of=None) -> np.ndarray:
    """Unmasked attention of every stream over its own K,V (attention.py:395-404, :1051-1058); kv_of maps a stream
    to the K/V stream it reads (cross attention of the compose variant)."""
    p = _empty(n_streams, heads)
    for s in range(n_streams):
        for h in range(heads):
            _add(p, s, h, s if kv_of is None else kv_of(s), 1.0)
    return p


def tca_plan(n_edits: int, heads: int, method: str, cg, src_id, tgt_id, kind: str = "edit",
             prefix: bool = False) -> np.ndarray:
    """Temporal_contextal_attention (attention.py:1043-1091) / _bg (:1284-1324) for n_edits edits of 4 streams.

    src_id(e) / tgt_id(e): bit-vector ids of edit e's fg_ref_mask (keys) and fg_retain_mask (rows); for kind='bg'
    src_id is the object mask (allowed keys = NOT object for every row) and tgt_id is unused.
    method 'mmsa': out = O_ref;  'tca': out = cg*O_ref + (1-cg)*O_self.
    prefix=True: the caller hands the kernel K,V whose ref-stream rows are sorted "source keys first"
    (kv_sort_index), so key masks become prefix lengths (FF_PASS_KEY_PREFIX): no per-element mask work, all-masked
    K/V tiles are skipped.
    """
    pf = FF_PASS_KEY_PREFIX if prefix else 0
    if method not in ("tca", "mmsa"):
        raise ValueError(f"method must be 'tca' or 'mmsa', got {method!r}")
    if kind not in ("edit", "bg"):
        raise ValueError(kind)
    p = _empty(4 * n_edits, heads)
    w_ref = 1.0 if method == "mmsa" else float(cg)
    w_self = 0.0 if method == "mmsa" else 1.0 - float(cg)
    for e in range(n_edits):
        for sl in range(4):
            s = 4 * e + sl
            r = 4 * e + (1 if sl < 2 else 3)                       # KV source: [u_r, u_r, c_r, c_r]
            for h in range(heads):
                masked = q0_masked(heads, sl, h)
                if not masked and r == s:
                    # the ref pass and the self pass coincide: cg*O + (1-cg)*O = O
                    _add(p, s, h, s, 1.0 if method == "mmsa" else w_ref + w_self)
                    continue
                if masked and kind == "edit":
                    # allowed(q,k) = (tgt[q] == src[k]) = src[k] ^ 1 ^ tgt[q]
                    _add(p, s, h, r, w_ref, key_mask=src_id(e), row_mask=tgt_id(e),
                         flags=FF_PASS_KEY_INVERT | FF_PASS_ROW_XOR | pf)
                elif masked:
                    _add(p, s, h, r, w_ref, key_mask=src_id(e), flags=FF_PASS_KEY_INVERT | pf)
                else:
                    _add(p, s, h, r, w_ref)
                if method == "tca":
                    _add(p, s, h, s, w_self)
    return p


def compose_plan(n_src: int, heads: int, method: str, cg, src_ids, tgt_ids, prefix: bool = False) -> np.ndarray:
    """Temporal_contextal_attention_compose (attention.py:1092-1140): streams [u_e, r_1..r_N, c_e]; the two edit
    streams get sum_i tgt_i(q) * softmax_{k in src_i}(q K_{r_i}) V_{r_i}; ref streams plain self-attention.  No Q0."""
    if method not in ("tca", "mmsa"):
        raise ValueError(method)
    B = n_src + 2
    p = _empty(B, heads)
    for s in range(B):
        for h in range(heads):
            if s in (0, B - 1):
                w_new = 1.0 if method == "mmsa" else float(cg)
                for i in range(n_src):
                    _add(p, s, h, 1 + i, w_new, key_mask=src_ids[i], row_mask=tgt_ids[i],
                         flags=FF_PASS_ROW_WEIGHT | (FF_PASS_KEY_PREFIX if prefix else 0))
                if method == "tca":
                    _add(p, s, h, s, 1.0 - float(cg))
            else:
                _add(p, s, h, s, 1.0)
    return p


def style_align_plan(n_edits: int, heads: int, src_id=None, prefix: bool = False, bg: bool = False) -> np.ndarray:
    """style_align_share_attention (attention.py:1142-1192): keys/values [self ; ref] under ONE softmax; with
    src_id (SDSA) the ref half is masked by fg_ref_mask on the Q0-masked (stream, head) pairs (:940-951).

    bg=True: style_align_share_attention_bg (:1193-1238) with prepare_sdsa_mask_for_bggen (:926-939), whose mask is
    1 - [ones ; obj]: on the Q0-masked pairs the WHOLE self half is masked out and the ref half admits only the keys
    OUTSIDE the object mask src_id(e) -- a single-segment pass over the ref stream with KEY_INVERT.  (Degenerate case: an
    object mask covering every token leaves no key; the reference then attends uniformly over [self ; ref] (quirk Q4),
    this plan uniformly over the ref keys.)"""
    p = _empty(4 * n_edits, heads)
    for e in range(n_edits):
        for sl in range(4):
            s = 4 * e + sl
            r = 4 * e + (1 if sl < 2 else 3)
            for h in range(heads):
                masked = src_id is not None and q0_masked(heads, sl, h)
                if bg and masked:
                    _add(p, s, h, r, 1.0, key_mask=src_id(e), flags=FF_PASS_KEY_INVERT | (FF_PASS_KEY_PREFIX if prefix else 0))
                    continue
                km2 = src_id(e) if masked else -1
                _add(p, s, h, s, 1.0, kv2=r, key_mask2=km2, flags=FF_PASS_KEY2_PREFIX if (prefix and km2 >= 0) else 0)
    return p


def kv_sort_index(key_bits, stream_masks) -> "torch.Tensor":
    """Row-gather index that sorts the keys of the masked K/V streams "set bits first" (stable), identity elsewhere.

    key_bits: bool/uint8 tensor [n_masks, S] (token-resolution masks, e.g. unpacked from ff_mask_downsample_pack);
    stream_masks: list of length n_kv_streams, entry = mask row that orders that stream's keys, or -1 (keep order).
    Returns int64 [n_kv_streams * S] flat row indices into the [n_kv_streams*S, C] view of K / V.  Softmax is
    permutation invariant over keys, so attention over (K[idx], V[idx]) with prefix masks equals attention over (K, V)
    with the bit-vector masks."""
    import torch
    n, S = key_bits.shape
    order = torch.sort(key_bits.to(torch.uint8), dim=1, descending=True, stable=True).indices      # [n_masks, S]
    ident = torch.arange(S, device=key_bits.device)
    rows = [order[m] if m >= 0 else ident for m in stream_masks]
    base = (torch.arange(len(stream_masks), device=key_bits.device) * S)[:, None]
    return (torch.stack(rows) + base).reshape(-1)


def algorithmic_flops(plan: np.ndarray, s_q: int, s_kv: int, d: int, popcount=None) -> float:
    """ALGORITHMIC FLOPs of one ff_attn_masked_kv launch: 4*d per (query, allowed key) pair of every
    pass -- QK^T + PV of exactly the pairs the softmax needs.  Padding (d -> 48/80/160), masked-out columns inside
    a partially masked tile and the hi/lo split of P are overhead, not work.  popcount[i] = set bits of bitmask row i."""
    total = 0.0
    for e in plan.reshape(-1):
        for ps in e["passes"][: int(e["n_pass"])]:
            flags = int(ps["flags"])
            n_r = int(popcount[int(ps["row_mask"])]) if ps["row_mask"] >= 0 else 0
            if ps["key_mask"] >= 0:
                n_k = int(popcount[int(ps["key_mask"])])
                c0 = s_kv - n_k if flags & FF_PASS_KEY_INVERT else n_k          # keys allowed for rows with rowbit 0
            else:
                c0 = s_kv
            if flags & FF_PASS_ROW_WEIGHT:
                pairs = n_r * c0
            elif flags & FF_PASS_ROW_XOR and ps["key_mask"] >= 0:
                pairs = (s_q - n_r) * c0 + n_r * (s_kv - c0)
            else:
                pairs = s_q * c0
            if ps["kv_stream2"] >= 0:
                if ps["key_mask2"] >= 0:
                    n2 = int(popcount[int(ps["key_mask2"])])
                    pairs += s_q * (s_kv - n2 if flags & FF_PASS_KEY2_INVERT else n2)
                else:
                    pairs += s_q * s_kv
            total += 4.0 * d * pairs
    return total


def describe(plan: np.ndarray) -> str:
    """Human-readable dump (debugging / DESIGN.md examples)."""
    out = []
    for s in range(plan.shape[0]):
        for h in range(plan.shape[1]):
            e = plan[s, h]
            ps = ", ".join(
                f"kv{int(q['kv_stream'])}" + (f"+kv{int(q['kv_stream2'])}" if q["kv_stream2"] >= 0 else "") +
                f"[k{int(q['key_mask'])} r{int(q['row_mask'])} f{int(q['flags'])}]*{float(q['weight']):.3f}"
                for q in e["passes"][: int(e["n_pass"])])
            out.append(f"s{s}h{h}: {ps}")
    return "\n".join(out)


__all__ = ["PASS_DTYPE", "PLAN_DTYPE", "q0_masked", "plain_plan", "tca_plan", "compose_plan", "style_align_plan",
           "kv_sort_index", "algorithmic_flops", "describe", "FF_PASS_KEY_INVERT", "FF_PASS_ROW_XOR", "FF_PASS_ROW_WEIGHT",
           "FF_PASS_KEY2_INVERT", "FF_PASS_KEY_PREFIX", "FF_PASS_KEY2_PREFIX"]
