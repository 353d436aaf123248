"""GPU parity of ff_attn_masked_kv (tcgen05 / TMEM / TMA) through the C ABI.

Tolerance: 2e-3 max-abs on the fp32 output (BASELINE.json north_star) against (i) the golden outputs of the
UNMODIFIED reference (tests/golden/attention.npz; inputs live on the bf16 grid so both sides see identical numbers)
and (ii) the CPU oracle on seeded inputs at sizes it finishes in seconds.  bf16 output adds one bf16 rounding."""
import numpy as np
import pytest
import torch

from freefine_b200 import plans
from oracle import cases, ff_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from freefine_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


P_OPERAND = "f16"


@pytest.fixture(autouse=True, params=["f16", "bf16x2"])
def p_operand(request):
    """Every parity test runs on both P.V operand paths: V staged as fp16 by ff_kv_gather_cast (one fp16 P operand) and
    V bf16 (hi+lo bf16 P pair)."""
    global P_OPERAND
    P_OPERAND = request.param
    return request.param


def _run(dev, q, k, v, plan, heads, scale, flat_masks=None, out_dtype=torch.float32, sort_streams=None):
    """sort_streams: per K/V stream the mask row that orders its keys (prefix mode) or -1."""
    from freefine_b200 import ops
    S = max(q.shape[1], k.shape[1])
    idx = None
    if sort_streams is not None:
        idx = plans.kv_sort_index(torch.from_numpy(np.stack([np.asarray(m) != 0 for m in flat_masks])), sort_streams)
    k, v = ops.kv_gather_cast(k.to(dev).bfloat16().contiguous(), v.to(dev).bfloat16().contiguous(), heads,
                              None if idx is None else idx.to(dev), p_operand=P_OPERAND)
    bm = pc = None
    if flat_masks is not None:
        words = ops.mask_words(S)
        arr = np.zeros((len(flat_masks), words), np.uint32)
        for i, m in enumerate(flat_masks):
            b = O.pack_bits(np.asarray(m) != 0)
            arr[i, : len(b)] = b
        bm = torch.from_numpy(arr.view(np.int32)).to(dev)
        pc = torch.tensor([int((np.asarray(m) != 0).sum()) for m in flat_masks], dtype=torch.int32, device=dev)
    pl = ops.to_device_bytes(plan, dev)
    out = ops.attn_masked_kv(q.to(dev).bfloat16(), k, v, pl, heads, scale, bm, pc, out_dtype=out_dtype)
    torch.cuda.synchronize()
    return out.float().cpu()


@pytest.mark.parametrize("name", list(cases.ATTN_CASES))
def test_tca_golden(dev, golden, name):
    g = golden["attention"]
    i = cases.attn_case_inputs(name)
    src, tgt = g[name + "/src_ds"], g[name + "/tgt_ds"]
    plan = plans.tca_plan(1, i["heads"], i["method"], i["cg"], lambda e: 0, lambda e: 1, kind=i["kind"])
    out = _run(dev, i["q"], i["k"], i["v"], plan, i["heads"], i["scale"], [src, tgt])
    ref = torch.from_numpy(g[name + "/out"])
    err = float((out - ref).abs().max())
    assert err < TOL, err
    ob = _run(dev, i["q"], i["k"], i["v"], plan, i["heads"], i["scale"], [src, tgt], out_dtype=torch.bfloat16)
    assert bool(((ob - ref).abs() <= TOL + 2.0 ** -8 * ref.abs()).all())
    # sorted-key (prefix) mode: same result from K,V with the ref streams' keys sorted "source first"
    pp = plans.tca_plan(1, i["heads"], i["method"], i["cg"], lambda e: 0, lambda e: 1, kind=i["kind"], prefix=True)
    op = _run(dev, i["q"], i["k"], i["v"], pp, i["heads"], i["scale"], [src, tgt], sort_streams=[-1, 0, -1, 0])
    assert float((op - ref).abs().max()) < TOL


def test_plain_cross_compose_style_golden(dev, golden):
    g = golden["attention"]
    T = lambda k: torch.from_numpy(g[k])
    sc = 8 ** -0.5
    out = _run(dev, T("plain/q"), T("plain/k"), T("plain/v"), plans.plain_plan(4, 8), 8, sc)
    assert float((out - T("plain/out")).abs().max()) < TOL
    # cross attention: 77 keys (ragged K/V tile), then the region blend kernel
    from freefine_b200 import ops
    reg = O.process_mask_before_attention(T("cross/region"), 64).numpy()
    hs = _run(dev, T("cross/q"), T("cross/k"), T("cross/v"), plans.plain_plan(4, 8), 8, sc)
    bits = torch.from_numpy(O.pack_bits(reg).view(np.int32))[None].to(dev)
    hs = ops.cross_region_blend(hs.to(dev), bits, torch.zeros(1, dtype=torch.int32, device=dev)).cpu()
    assert float((hs - T("cross/out")).abs().max()) < TOL
    srcs = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/srcs")]
    tgts = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/tgts")]
    for method, cg in (("tca", 0.3), ("mmsa", None)):
        plan = plans.compose_plan(2, 8, method, cg, [0, 1], [2, 3])
        out = _run(dev, T("compose/q"), T("compose/k"), T("compose/v"), plan, 8, sc, srcs + tgts)
        assert float((out - T(f"compose_{method}/out")).abs().max()) < TOL, method
    # compose cross-attention: q stream i reads prompt stream kv_of(i); c_e = sum_i tgt_i * Attn(q_c, prompt_i)
    q, k, v = T("cross_compose/q"), T("cross_compose/k"), T("cross_compose/v")
    p = plans._empty(4, 8)
    for s in range(3):
        for h in range(8):
            plans._add(p, s, h, s, 1.0)
    for h in range(8):
        for i in range(2):
            plans._add(p, 3, h, 3 + i, 1.0, row_mask=i, flags=plans.FF_PASS_ROW_WEIGHT)
    out = _run(dev, q, k, v, p, 8, sc, tgts)
    assert float((out - T("cross_compose/out")).abs().max()) < TOL
    src = O.process_mask_before_attention(T("style/src"), 64).numpy()
    out = _run(dev, T("style/q"), T("style/k"), T("style/v"), plans.style_align_plan(1, 8), 8, sc, [src])
    assert float((out - T("style_ssa/out")).abs().max()) < TOL
    out = _run(dev, T("style/q"), T("style/k"), T("style/v"), plans.style_align_plan(1, 8, lambda e: 0), 8, sc, [src])
    assert float((out - T("style_sdsa/out")).abs().max()) < TOL
    out = _run(dev, T("style/q"), T("style/k"), T("style/v"), plans.style_align_plan(1, 8, lambda e: 0, prefix=True), 8, sc,
               [src], sort_streams=[-1, 0, -1, 0])
    assert float((out - T("style_sdsa/out")).abs().max()) < TOL
    plan = plans.compose_plan(2, 8, "tca", 0.3, [0, 1], [2, 3], prefix=True)
    out = _run(dev, T("compose/q"), T("compose/k"), T("compose/v"), plan, 8, sc, srcs + tgts, sort_streams=[-1, 0, 1, -1])
    assert float((out - T("compose_tca/out")).abs().max()) < TOL


@pytest.mark.parametrize("S,d,res,method,kind", [(1024, 80, 256, "tca", "edit"), (1024, 40, 256, "tca", "edit"),
                                                  (576, 160, 192, "tca", "edit"), (144, 160, 96, "mmsa", "bg"),
                                                  (2304, 40, 384, "tca", "bg")])
def test_tca_vs_oracle_sd15_shapes(dev, S, d, res, method, kind):
    """SD1.5 layer shapes incl. the ragged 768^2 ones (576 = 4.5 tiles, 144, 2304 = 18 tiles); 2 batched edits."""
    heads, E = 8, 2
    q, k, v = cases.qkv(4 * E, S, heads * d, 500 + S + d)
    flat = []
    for e in range(E):
        flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 700 + e)), S).numpy())      # src
        flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(res, 800 + e)), S).numpy())      # tgt
    cg = 0.55
    refs = [O.tca(q[4 * e:4 * e + 4], k[4 * e:4 * e + 4], v[4 * e:4 * e + 4], heads, d ** -0.5, flat[2 * e], flat[2 * e + 1],
                  method, cg, kind=kind) for e in range(E)]
    for prefix in (False, True):
        plan = plans.tca_plan(E, heads, method, cg, lambda e: 2 * e, lambda e: 2 * e + 1, kind=kind, prefix=prefix)
        out = _run(dev, q, k, v, plan, heads, d ** -0.5, flat,
                   sort_streams=[2 * (s // 4) if s % 2 else -1 for s in range(4 * E)] if prefix else None)
        for e in range(E):
            err = float((out[4 * e:4 * e + 4] - refs[e]).abs().max())
            assert err < TOL, (prefix, e, err)


def test_peaked_softmax_and_large_logits(dev):
    """|logits| ~ 30: exercises the lazy-rescale path (running max grows by > 2^8 across tiles)."""
    heads, d, S = 8, 40, 1024
    q, k, v = cases.qkv(4, S, heads * d, 901, logit_scale=12.0)
    # make later keys systematically larger so the row max keeps growing tile after tile
    k = (k * torch.linspace(0.2, 2.0, S)[None, :, None]).bfloat16().float()
    src = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(256, 902)), S).numpy()
    tgt = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(256, 903)), S).numpy()
    plan = plans.tca_plan(1, heads, "tca", 0.5, lambda e: 0, lambda e: 1)
    out = _run(dev, q, k, v, plan, heads, d ** -0.5, [src, tgt])
    ref = O.tca(q, k, v, heads, d ** -0.5, src, tgt, "tca", 0.5)
    assert float((out - ref).abs().max()) < TOL


def test_full_size_properties(dev):
    """BASELINE size S=4096, d=40, 8 batched edits (32 streams): size-independent properties instead of an oracle run.
    (1) rows of V constant per head  => output equals that constant whatever the masks (softmax weights sum to 1);
    (2) linearity in V;  (3) batched edits equal the same edit run alone (bit-exact, same kernel)."""
    heads, d, S, E = 8, 40, 4096, 8
    q, k, v = cases.qkv(4 * E, S, heads * d, 1001)
    flat = []
    for e in range(E):
        flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(512, 1100 + e, (0.08, 0.2))), S).numpy())
        flat.append(O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(512, 1200 + e, (0.08, 0.2))), S).numpy())
    plan = plans.tca_plan(E, heads, "tca", 0.4, lambda e: 2 * e, lambda e: 2 * e + 1, prefix=True)
    sort = [2 * (s // 4) if s % 2 else -1 for s in range(4 * E)]
    const = torch.randn(4 * E, 1, heads * d, generator=torch.Generator().manual_seed(5)).bfloat16().float()
    # per edit the ref streams (1,3) provide V for the ref passes, the stream itself for the self pass: make V
    # constant over tokens AND equal across the 4 streams of an edit so every pass returns the same constant
    vc = const[::4].repeat_interleave(4, 0).expand(4 * E, S, heads * d).contiguous()
    out_c = _run(dev, q, k, vc, plan, heads, d ** -0.5, flat, sort_streams=sort)
    # (the tensor core accumulates 4096 products per row in fp32 with truncation: a few 1e-5 relative)
    assert float((out_c - vc).abs().max()) < 2e-4 * float(vc.abs().max())
    out1 = _run(dev, q, k, v, plan, heads, d ** -0.5, flat, sort_streams=sort)
    out2 = _run(dev, q, k, (2 * v), plan, heads, d ** -0.5, flat, sort_streams=sort)
    out_bits = _run(dev, q, k, v, plans.tca_plan(E, heads, "tca", 0.4, lambda e: 2 * e, lambda e: 2 * e + 1), heads, d ** -0.5, flat)
    # bit-vector mode == sorted-key mode up to the rounding of P (key order differs): fp16 P ~1e-3, hi+lo bf16 ~1e-5
    assert float((out_bits - out1).abs().max()) < (TOL if P_OPERAND == "f16" else 2e-4)
    assert float((out2 - 2 * out1).abs().max()) < 1e-4
    assert bool(torch.isfinite(out1).all())
    e = 3
    p1 = plans.tca_plan(1, heads, "tca", 0.4, lambda _: 0, lambda _: 1, prefix=True)
    alone = _run(dev, q[4 * e:4 * e + 4], k[4 * e:4 * e + 4], v[4 * e:4 * e + 4], p1, heads, d ** -0.5, flat[2 * e:2 * e + 2],
                 sort_streams=[-1, 0, -1, 0])
    assert torch.equal(alone, out1[4 * e:4 * e + 4])
    # one head-slice of one stream against the oracle (seconds on CPU): masked head 0 of the cond-edit stream
    s, h = 4 * e + 2, 0
    qh = q[s, :, :d][None]; kr = k[4 * e + 3, :, :d][None]; vr = v[4 * e + 3, :, :d][None]
    ks = k[s, :, :d][None]; vs = v[s, :, :d][None]
    srcb = torch.from_numpy(flat[2 * e] != 0); tgtb = torch.from_numpy(flat[2 * e + 1] != 0)
    allowed = torch.where(tgtb[:, None], srcb[None, :], ~srcb[None, :])
    ref = 0.4 * O._softmax_av(qh[0], kr[0], vr[0], d ** -0.5, allowed) + 0.6 * O._softmax_av(qh[0], ks[0], vs[0], d ** -0.5)
    assert float((out1[s, :, :d] - ref).abs().max()) < TOL


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_plans_vs_plan_interpreter(dev, seed):
    """Fuzz: random shapes (ragged S_q / S_kv, 1..5 K/V tiles, every head-dim instantiation), random bit-vector masks
    incl. empty and full ones, random multi-pass plans (1-3 passes, second K/V segment, KEY_INVERT / ROW_XOR /
    ROW_WEIGHT) against the literal CPU reading of the plan semantics (tests/plan_interp.py).  Exercises single-tile
    passes, passes in which only one of the two softmax warpgroups gets a tile, rows that read nothing from a tile."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from plan_interp import run_plan
    from freefine_b200 import ops
    if seed % 3 == 0:
        # stale allocator contents must not matter: poison the free blocks the next allocations will reuse
        junk = [torch.full((n,), float("nan"), device=dev) for n in (1 << 12, 1 << 15, 1 << 18, 1 << 21, 1 << 23)]
        junk += [torch.full((n,), 0xFF, dtype=torch.uint8, device=dev) for n in (1 << 10, 1 << 14, 1 << 17)]
        torch.cuda.synchronize()
        del junk
    rng = np.random.default_rng(9000 + seed)
    B = int(rng.integers(1, 4))
    heads = int(rng.choice([1, 2, 4]))
    d = int(rng.choice([8, 16, 24, 40, 64, 80, 96, 160]))
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
                kv = int(rng.integers(0, B))
                km = int(rng.integers(-1, 4))
                rm = int(rng.integers(-1, 4))
                flags = 0
                if km >= 0 and rng.random() < 0.4:
                    flags |= plans.FF_PASS_KEY_INVERT
                if km >= 0 and rm >= 0 and rng.random() < 0.5:
                    flags |= plans.FF_PASS_ROW_XOR
                if rm >= 0 and rng.random() < 0.3:
                    flags |= plans.FF_PASS_ROW_WEIGHT
                kv2, km2 = -1, -1
                # second segment under the same softmax: every key of another stream (a segment WITHOUT a key mask admits
                # every key whatever the flags say, so ROW_XOR passes stay single-segment here)
                if not (flags & plans.FF_PASS_ROW_XOR) and rng.random() < 0.25:
                    kv2 = int(rng.integers(0, B))
                plans._add(plan, s_, h, kv, float(rng.uniform(0.1, 1.0)), key_mask=km, row_mask=rm, flags=flags, kv2=kv2,
                           key_mask2=km2)
    words = ops.mask_words(S)
    arr = np.zeros((len(flat), words), np.uint32)
    for i, m in enumerate(flat):
        b = O.pack_bits(m != 0)
        arr[i, : len(b)] = b
    # (popcounts are over the first S_kv tokens: what the kernel's empty-set rule looks at)
    flat_kv = [m.copy() for m in flat]
    ref = run_plan(q, k, v, plan, heads, d ** -0.5, arr)
    bm = torch.from_numpy(arr.view(np.int32)).to(dev)
    pc = torch.tensor([int(m[:Skv].sum()) for m in flat_kv], dtype=torch.int32, device=dev)
    kk, vv = ops.kv_gather_cast(k.to(dev).bfloat16().contiguous(), v.to(dev).bfloat16().contiguous(), heads, None,
                                p_operand=P_OPERAND)
    out = ops.attn_masked_kv(q.to(dev).bfloat16(), kk, vv, ops.to_device_bytes(plan, dev), heads, d ** -0.5, bm, pc,
                             out_dtype=torch.float32)
    torch.cuda.synchronize()
    err = float((out.cpu() - ref).abs().max())
    assert err < TOL * 3.0, (err, B, heads, d, Sq, Skv)       # up to 3 passes of weight <= 1 each


def test_error_paths(dev):
    from freefine_b200 import ops
    q = torch.zeros(4, 64, 64, device=dev, dtype=torch.bfloat16)
    pl = ops.to_device_bytes(plans.plain_plan(4, 8), dev)
    with pytest.raises(TypeError):
        ops.attn_masked_kv(q.float(), q, q, pl, 8, 1.0)
    with pytest.raises(ValueError):
        ops.attn_masked_kv(q, q, q, pl[:10], 8, 1.0)
    with pytest.raises(RuntimeError):            # head_dim = 4 is not a multiple of 8 -> library error, no fallback
        ops.attn_masked_kv(q, q, q, ops.to_device_bytes(plans.plain_plan(4, 16), dev), 16, 1.0)
    with pytest.raises(RuntimeError):
        ops.attn_masked_kv(q.cpu(), q, q, pl, 8, 1.0)


def test_style_align_bg_golden(dev, golden):
    """style_align_share_attention_bg (attention.py:1193-1238): 'ssa' = [self ; ref] under one softmax; 'sdsa' with
    prepare_sdsa_mask_for_bggen (:926-939): self half masked out, ref keys outside the object mask, on the Q0 pairs."""
    g = golden["attention_bg"]
    T = lambda k: torch.from_numpy(g[k])
    sc = 8 ** -0.5
    obj = O.process_mask_before_attention(T("style_bg/obj"), 64).numpy()
    q, k, v = T("style_bg/q"), T("style_bg/k"), T("style_bg/v")
    out = _run(dev, q, k, v, plans.style_align_plan(1, 8, None, bg=True), 8, sc, [obj])
    assert float((out - T("style_bg_ssa/out")).abs().max()) < TOL
    out = _run(dev, q, k, v, plans.style_align_plan(1, 8, lambda e: 0, bg=True), 8, sc, [obj])
    assert float((out - T("style_bg_sdsa/out")).abs().max()) < TOL
    out = _run(dev, q, k, v, plans.style_align_plan(1, 8, lambda e: 0, prefix=True, bg=True), 8, sc, [obj], sort_streams=[-1, 0, -1, 0])
    assert float((out - T("style_bg_sdsa/out")).abs().max()) < TOL


def test_s9216_head_slices_vs_oracle(dev):
    """BASELINE.json configs[4]: 768^2 -> 96x96 latents, S = 9216, d = 40 (the ragged-free 144-tile case).  One edit
    (4 streams x 8 heads) on the GPU; a masked head of the cond-edit stream, an unmasked head of the uncond-edit stream and a
    masked head of a ref stream against the CPU oracle (one head slice each: seconds on the CPU)."""
    heads, d, S = 8, 40, 9216
    q, k, v = cases.qkv(4, S, heads * d, 1301)
    src = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(768, 1302, (0.08, 0.2))), S).numpy()
    tgt = O.process_mask_before_attention(torch.from_numpy(cases.blob_mask(768, 1303, (0.08, 0.2))), S).numpy()
    cg = 0.35
    plan = plans.tca_plan(1, heads, "tca", cg, lambda e: 0, lambda e: 1, prefix=True)
    out = _run(dev, q, k, v, plan, heads, d ** -0.5, [src, tgt], sort_streams=[-1, 0, -1, 0])
    srcb, tgtb = torch.from_numpy(src != 0), torch.from_numpy(tgt != 0)
    allowed = torch.where(tgtb[:, None], srcb[None, :], ~srcb[None, :])
    sl = lambda t, s, h: t[s, :, h * d:(h + 1) * d]
    for s, h in ((2, 0), (0, 1), (3, 2)):
        r = 1 if s < 2 else 3
        masked = plans.q0_masked(heads, s, h)
        ref_pass = O._softmax_av(sl(q, s, h), sl(k, r, h), sl(v, r, h), d ** -0.5, allowed if masked else None)
        self_pass = O._softmax_av(sl(q, s, h), sl(k, s, h), sl(v, s, h), d ** -0.5)
        ref = cg * ref_pass + (1 - cg) * self_pass
        err = float((sl(out, s, h) - ref).abs().max())
        assert err < TOL, (s, h, err)


# ---- ff_attn_plain_smallkv: plain attention over <= 128 keys (text cross-attention, 8 x 8 self-attention) -------------------
def _smallkv(dev, q, k, v, heads, scale, out_dtype=torch.float32):
    from freefine_b200 import ops
    out = ops.attn_plain_smallkv(q.to(dev).bfloat16().contiguous(), k.to(dev).bfloat16().contiguous(),
                                 v.to(dev).bfloat16().contiguous(), heads, scale, out_dtype=out_dtype)
    torch.cuda.synchronize()
    return out.float().cpu()


def test_smallkv_golden(dev, golden):
    """The reference's own plain / 77-key cross-attention outputs (tests/golden/attention.npz) through the short-key kernel."""
    if P_OPERAND != "f16":
        pytest.skip("ff_attn_plain_smallkv has one operand path (fp16 P.V)")
    g = golden["attention"]
    T = lambda k: torch.from_numpy(g[k])
    sc = 8 ** -0.5
    assert T("cross/k").shape[1] == 77
    hs = _smallkv(dev, T("cross/q"), T("cross/k"), T("cross/v"), 8, sc)
    assert float((hs - _run(dev, T("cross/q"), T("cross/k"), T("cross/v"), plans.plain_plan(4, 8), 8, sc)).abs().max()) < 1e-3
    from freefine_b200 import ops
    reg = O.process_mask_before_attention(T("cross/region"), 64).numpy()
    bits = torch.from_numpy(O.pack_bits(reg).view(np.int32))[None].to(dev)
    hs = ops.cross_region_blend(hs.to(dev), bits, torch.zeros(1, dtype=torch.int32, device=dev)).cpu()
    assert float((hs - T("cross/out")).abs().max()) < TOL
    if T("plain/k").shape[1] <= 128:
        out = _smallkv(dev, T("plain/q"), T("plain/k"), T("plain/v"), 8, sc)
        assert float((out - T("plain/out")).abs().max()) < TOL


@pytest.mark.parametrize("B,heads,d,s_q,s_kv", [(2, 8, 40, 4096, 77), (2, 8, 80, 1024, 77), (3, 8, 160, 256, 77), (2, 8, 160, 64, 64),
                                                (1, 2, 40, 200, 77), (1, 2, 80, 70, 128), (2, 2, 160, 17, 1), (1, 3, 40, 333, 81),
                                                (1, 2, 8, 100, 5),
                                                # two key blocks under one online softmax: the plain 16 x 16 self-attention (256 keys), ragged 129 / 200
                                                (2, 8, 160, 256, 256), (1, 2, 40, 100, 200), (1, 2, 80, 130, 129), (1, 2, 160, 77, 255)])
def test_smallkv_vs_oracle(dev, B, heads, d, s_q, s_kv):
    """Seeded inputs on the bf16 grid against the CPU oracle's plain attention: SD1.5 shapes (64^2 / 32^2 / 16^2 cross-attention,
    8^2 / 16^2 self-attention), ragged row counts (not multiples of 16 / 64 / 256), 1 / 81 / 128 / 129 / 256 keys, f32 and bf16 outputs."""
    if P_OPERAND != "f16":
        pytest.skip("ff_attn_plain_smallkv has one operand path (fp16 P.V)")
    g = torch.Generator().manual_seed(1000 * d + s_q + s_kv)
    C = heads * d
    q = torch.randn(B, s_q, C, generator=g).bfloat16().float()
    k = torch.randn(B, s_kv, C, generator=g).bfloat16().float()
    v = (2.0 * torch.randn(B, s_kv, C, generator=g)).bfloat16().float()
    sc = d ** -0.5
    want = O.plain_attention(q, k, v, heads, sc)
    got = _smallkv(dev, q, k, v, heads, sc)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < TOL
    got16 = _smallkv(dev, q, k, v, heads, sc, out_dtype=torch.bfloat16)
    assert float((got16 - want).abs().max()) < TOL + 2.0 ** -8 * float(want.abs().max())
    # peaked rows: one dominant key per query (large logits), the softmax must not overflow or lose the row
    k2 = k.clone()
    k2[:, 0] = 30.0 * q[:, 0, :].sign()
    want2 = O.plain_attention(q, k2, v, heads, sc)
    got2 = _smallkv(dev, q, k2, v, heads, sc)
    assert torch.isfinite(got2).all() and float((got2 - want2).abs().max()) < 2 * TOL


def test_smallkv_error_paths(dev):
    from freefine_b200 import ops
    q = torch.zeros(1, 32, 80, device=dev, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="s_kv"):
        ops.attn_plain_smallkv(q, torch.zeros(1, 257, 80, device=dev, dtype=torch.bfloat16), torch.zeros(1, 257, 80, device=dev, dtype=torch.bfloat16), 2, 0.1)
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.attn_plain_smallkv(q, torch.zeros(1, 8, 80, device=dev, dtype=torch.bfloat16), torch.zeros(1, 8, 80, device=dev, dtype=torch.bfloat16), 5, 0.1)
    with pytest.raises(ValueError):
        ops.attn_plain_smallkv(q, torch.zeros(2, 8, 80, device=dev, dtype=torch.bfloat16), torch.zeros(2, 8, 80, device=dev, dtype=torch.bfloat16), 2, 0.1)
