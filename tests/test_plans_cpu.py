"""Host logic (freefine_b200/plans.py) on CPU: the plans, read literally (tests/plan_interp.py), reproduce the
reference outputs stored in tests/golden/attention.npz and the oracle.  No GPU, no CUDA library call."""
import numpy as np
import pytest
import torch

from freefine_b200 import plans
from oracle import cases, ff_oracle as O
from plan_interp import run_plan


def _bits(*flat_masks):
    return np.stack([O.pack_bits(m) for m in flat_masks])


@pytest.mark.parametrize("name", [n for n in cases.ATTN_CASES if cases.ATTN_CASES[n]["S"] <= 256 and cases.ATTN_CASES[n]["d"] <= 40])
def test_tca_plan_matches_reference(golden, name):
    g = golden["attention"]
    i = cases.attn_case_inputs(name)
    src, tgt = g[name + "/src_ds"], g[name + "/tgt_ds"]
    bm = _bits(src, tgt)
    plan = plans.tca_plan(1, i["heads"], i["method"], i["cg"], lambda e: 0, lambda e: 1, kind=i["kind"])
    out = run_plan(i["q"], i["k"], i["v"], plan, i["heads"], i["scale"], bm)
    err = float((out - torch.from_numpy(g[name + "/out"])).abs().max())
    assert err < 5e-5, err


def test_q0_rule():
    # heads=8: even heads of EVERY stream are masked; heads=2: stream parity matters too
    assert [plans.q0_masked(8, s, h) for s in range(4) for h in range(8)] == [h % 2 == 0 for s in range(4) for h in range(8)]
    assert [plans.q0_masked(2, s, h) for s in range(4) for h in range(2)] == [((2 * s + h) % 4) in (0, 2) for s in range(4) for h in range(2)]
    assert plans.q0_masked(8, 1, 0) == O.q0_masked(8, 1, 0)


def test_tca_plan_structure():
    p = plans.tca_plan(2, 8, "tca", 0.25, lambda e: 2 * e, lambda e: 2 * e + 1)
    assert p.shape == (8, 8)
    # ref streams, odd heads: ref pass and self pass coincide -> one pass of weight 1
    assert int(p[1, 1]["n_pass"]) == 1 and float(p[1, 1]["passes"][0]["weight"]) == 1.0
    # edit stream of edit 1, even head: masked ref pass (keys = mask 2, rows = mask 3) + self pass
    e = p[4, 0]
    assert int(e["n_pass"]) == 2
    assert (int(e["passes"][0]["kv_stream"]), int(e["passes"][0]["key_mask"]), int(e["passes"][0]["row_mask"])) == (5, 2, 3)
    assert int(e["passes"][0]["flags"]) == plans.FF_PASS_KEY_INVERT | plans.FF_PASS_ROW_XOR
    assert int(e["passes"][1]["kv_stream"]) == 4 and abs(float(e["passes"][1]["weight"]) - 0.75) < 1e-7
    # cond streams read c_r (stream 3 of their edit)
    assert int(p[6, 3]["passes"][0]["kv_stream"]) == 7
    with pytest.raises(ValueError):
        plans.tca_plan(1, 8, "ssa", 0.5, lambda e: 0, lambda e: 1)
    # dense head-pass count of one 'tca' layer call: 32 ref passes + 24 distinct self passes
    p1 = plans.tca_plan(1, 8, "tca", 0.5, lambda e: 0, lambda e: 1)
    assert int(p1["n_pass"].sum()) == 56
    assert int(plans.tca_plan(1, 8, "mmsa", None, lambda e: 0, lambda e: 1)["n_pass"].sum()) == 32


def test_compose_and_style_plans(golden):
    g = golden["attention"]
    T = lambda k: torch.from_numpy(g[k])
    sc = 8 ** -0.5
    srcs = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/srcs")]
    tgts = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/tgts")]
    bm = _bits(*srcs, *tgts)
    for method, cg in (("tca", 0.3), ("mmsa", None)):
        plan = plans.compose_plan(2, 8, method, cg, [0, 1], [2, 3])
        out = run_plan(T("compose/q"), T("compose/k"), T("compose/v"), plan, 8, sc, bm)
        assert float((out - T(f"compose_{method}/out")).abs().max()) < 2e-5
    src = O.process_mask_before_attention(T("style/src"), 64).numpy()
    bm = _bits(src)
    out = run_plan(T("style/q"), T("style/k"), T("style/v"), plans.style_align_plan(1, 8), 8, sc, bm)
    assert float((out - T("style_ssa/out")).abs().max()) < 2e-5
    out = run_plan(T("style/q"), T("style/k"), T("style/v"), plans.style_align_plan(1, 8, lambda e: 0), 8, sc, bm)
    assert float((out - T("style_sdsa/out")).abs().max()) < 2e-5
    out = run_plan(T("plain/q"), T("plain/k"), T("plain/v"), plans.plain_plan(4, 8), 8, sc)
    assert float((out - T("plain/out")).abs().max()) < 2e-5


def test_batched_edits_equal_sequential():
    """E edits in one launch == E sequential 4-stream runs (the reference cannot batch: attention.py:1034)."""
    q, k, v = cases.qkv(8, 64, 64, 77)
    masks = [cases.blob_mask(64, 300 + i) for i in range(4)]
    flat = [O.process_mask_before_attention(torch.from_numpy(m), 64).numpy() for m in masks]
    bm = _bits(*flat)
    plan = plans.tca_plan(2, 8, "tca", 0.6, lambda e: 2 * e, lambda e: 2 * e + 1)
    out = run_plan(q, k, v, plan, 8, 8 ** -0.5, bm)
    for e in range(2):
        ref = O.tca(q[4 * e:4 * e + 4], k[4 * e:4 * e + 4], v[4 * e:4 * e + 4], 8, 8 ** -0.5, flat[2 * e], flat[2 * e + 1], "tca", 0.6)
        assert float((out[4 * e:4 * e + 4] - ref).abs().max()) < 2e-5


def test_prefix_mode_sorted_keys_equals_bitmask_mode(golden):
    """FF_PASS_KEY_PREFIX + keys sorted "mask bits first" (plans.kv_sort_index) == the bit-vector plan (softmax is
    permutation invariant over keys)."""
    g = golden["attention"]
    for name in ("tca_h8_s256", "bg_tca_h8_s256", "tca_h8_s64_onekey", "tca_h8_s64_empty", "tca_h8_s64_full"):
        i = cases.attn_case_inputs(name)
        src, tgt = g[name + "/src_ds"], g[name + "/tgt_ds"]
        bm = _bits(src, tgt)
        plan = plans.tca_plan(1, i["heads"], i["method"], i["cg"], lambda e: 0, lambda e: 1, kind=i["kind"], prefix=True)
        idx = plans.kv_sort_index(torch.from_numpy(np.stack([src, tgt]) != 0), [-1, 0, -1, 0])
        S, C = i["k"].shape[1:]
        ks = i["k"].reshape(-1, C)[idx].reshape(4, S, C)
        vs = i["v"].reshape(-1, C)[idx].reshape(4, S, C)
        out = run_plan(i["q"], ks, vs, plan, i["heads"], i["scale"], bm)
        assert float((out - torch.from_numpy(g[name + "/out"])).abs().max()) < 5e-5, name
    # SDSA / compose variants
    T = lambda k: torch.from_numpy(g[k])
    src = O.process_mask_before_attention(T("style/src"), 64).numpy()
    idx = plans.kv_sort_index(torch.from_numpy(src[None] != 0), [-1, 0, -1, 0])
    k, v = T("style/k"), T("style/v")
    ks, vs = k.reshape(-1, 64)[idx].reshape(k.shape), v.reshape(-1, 64)[idx].reshape(v.shape)
    out = run_plan(T("style/q"), ks, vs, plans.style_align_plan(1, 8, lambda e: 0, prefix=True), 8, 8 ** -0.5, _bits(src))
    assert float((out - T("style_sdsa/out")).abs().max()) < 2e-5
    srcs = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/srcs")]
    tgts = [O.process_mask_before_attention(m, 64).numpy() for m in T("compose/tgts")]
    idx = plans.kv_sort_index(torch.from_numpy(np.stack(srcs) != 0), [-1, 0, 1, -1])
    k, v = T("compose/k"), T("compose/v")
    ks, vs = k.reshape(-1, 64)[idx].reshape(k.shape), v.reshape(-1, 64)[idx].reshape(v.shape)
    out = run_plan(T("compose/q"), ks, vs, plans.compose_plan(2, 8, "tca", 0.3, [0, 1], [2, 3], prefix=True), 8, 8 ** -0.5,
                   _bits(*srcs, *tgts))
    assert float((out - T("compose_tca/out")).abs().max()) < 2e-5


def test_composition_source_count_is_checked_up_front():
    """N sources need N + 1 passes per head (FF_MAX_PASS = 4): the entry point refuses N = 4 before any UNet call."""
    from freefine_b200._lib import FF_MAX_PASS
    from freefine_b200.pipeline import FreeFinePipeline
    import pytest
    with pytest.raises(ValueError, match="FF_MAX_PASS"):
        FreeFinePipeline.Details_Preserving_regeneration_compose(object(), None, [None], ["a"] * FF_MAX_PASS,
                                                                 [[0]] * FF_MAX_PASS, [[0]] * (FF_MAX_PASS + 1), None)
