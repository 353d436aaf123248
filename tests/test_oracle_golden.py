"""Pins oracle/ff_oracle.py (the CPU restatement) against the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cases, ff_oracle as O


def _chk(i, g, name):
    cs = np.array([float(i[t].double().sum()) for t in "qkv"] + [float(i[t].double().abs().sum()) for t in "qkv"])
    np.testing.assert_allclose(cs, g[name + "/qkv_checksum"], rtol=1e-12, err_msg="seeded inputs drifted")


@pytest.mark.parametrize("name", list(cases.ATTN_CASES))
def test_tca_matches_reference(golden, name):
    g = golden["attention"]
    i = cases.attn_case_inputs(name)
    _chk(i, g, name)
    src = O.process_mask_before_attention(i["src"], i["S"])
    tgt = O.process_mask_before_attention(i["tgt"], i["S"])
    # integer work: bit exact
    assert np.array_equal(src.numpy(), g[name + "/src_ds"])
    assert np.array_equal(tgt.numpy(), g[name + "/tgt_ds"])
    out = O.tca(i["q"], i["k"], i["v"], i["heads"], i["scale"], src, tgt, i["method"], i["cg"], kind=i["kind"])
    err = float((out - torch.from_numpy(g[name + "/out"])).abs().max())
    assert err < 5e-5, err


def test_plain_cross_compose_style(golden):
    g = golden["attention"]
    T = lambda k: torch.from_numpy(g[k])
    sc = 8 ** -0.5
    out = O.plain_attention(T("plain/q"), T("plain/k"), T("plain/v"), 8, sc)
    assert float((out - T("plain/out")).abs().max()) < 2e-5
    reg = O.process_mask_before_attention(T("cross/region"), 64)
    out = O.cross_local(T("cross/q"), T("cross/k"), T("cross/v"), 8, sc, reg)
    assert float((out - T("cross/out")).abs().max()) < 2e-5
    srcs = [O.process_mask_before_attention(m, 64) for m in T("compose/srcs")]
    tgts = [O.process_mask_before_attention(m, 64) for m in T("compose/tgts")]
    for method, cg in (("tca", 0.3), ("mmsa", None)):
        out = O.tca_compose(T("compose/q"), T("compose/k"), T("compose/v"), 8, sc, srcs, tgts, method, cg)
        assert float((out - T(f"compose_{method}/out")).abs().max()) < 2e-5
    out = O.cross_local_compose(T("cross_compose/q"), T("cross_compose/k"), T("cross_compose/v"), 8, sc, tgts, 2)
    assert float((out - T("cross_compose/out")).abs().max()) < 2e-5
    src = O.process_mask_before_attention(T("style/src"), 64)
    out = O.style_align(T("style/q"), T("style/k"), T("style/v"), 8, sc, None)
    assert float((out - T("style_ssa/out")).abs().max()) < 2e-5
    out = O.style_align(T("style/q"), T("style/k"), T("style/v"), 8, sc, src)
    assert float((out - T("style_sdsa/out")).abs().max()) < 2e-5


def test_steps_bit_exact(golden):
    g = golden["steps"]
    al = O.make_alphas_cumprod()
    for n_steps, eta, seed in ((50, 1.0, 1), (50, 0.0, 2), (10, 1.0, 3), (10, 0.5, 4)):
        eps4, x, noise, cfg_mask, var_mask = cases.step_case_inputs(seed)
        assert np.array_equal(eps4.numpy(), g[f"seed{seed}/eps4"])
        ts = O.timesteps_for(n_steps)
        for t in (int(ts[0]), int(ts[len(ts) // 2]), int(ts[-1])):
            key = f"n{n_steps}_eta{eta}_t{t}"
            eu, ec = eps4.chunk(2)
            eps = O.cfg_local(eu, ec, 7.5, cfg_mask)
            assert np.array_equal(eps.numpy(), g[key + "/cfg"])
            xp, x0 = O.ctrl_step(eps, t, x, var_mask, eta, noise, al, n_steps)
            assert np.array_equal(xp.numpy(), g[key + "/x_prev"]), key
            assert np.array_equal(x0.numpy(), g[key + "/pred_x0"]), key
            xn, x0i = O.inv_step(eps4[:2], t, x, al, n_steps)
            assert np.array_equal(xn.numpy(), g[key + "/x_next"]), key
            assert np.array_equal(x0i.numpy(), g[key + "/inv_x0"]), key
    lp = np.array([[O.linear_param(i, 35, 50, 50, 0.0) for i in range(35, 51)],
                   [O.linear_param(i, 0, 10, 16, 0.5) for i in range(0, 16)]])
    assert np.array_equal(lp, g["linear_param"])


def test_warp(golden):
    g = golden["warp"]
    for seed in (1, 2, 3):
        src, bg, mask, M = cases.warp_case_inputs(seed)
        assert np.array_equal(src, g[f"s{seed}/src"])
        H, W = src.shape[-2:]
        theta = O.param2theta(M, W, H)
        assert np.array_equal(theta, g[f"s{seed}/theta"])
        wb = O.warp_affine(src, theta, (W, H), "bilinear")
        assert np.abs(wb - g[f"s{seed}/warp_bilinear"].reshape(wb.shape)).max() < 1e-5
        wn = O.warp_affine(mask.astype(np.float32)[None, None], theta, (W, H), "nearest")
        assert np.array_equal(wn.reshape(H, W), g[f"s{seed}/warp_nearest"].reshape(H, W))     # index work: exact
        bl, _ = O.warp_blend(src, theta, mask, bg)
        assert np.abs(bl - g[f"s{seed}/blend"]).max() < 1e-5


def test_mask_prep_bit_exact(golden):
    g = golden["masks"]
    for auto in (False, True):
        for red in (False, True):
            r = O.prepare_various_mask(g["shifted"], g["ori"], g["draw"], 128, 128, 16, 16, use_auto_draw=auto,
                                       cons_area=g["cons"], reduce_inp_artifacts=red)
            for nm, t in zip(("fg", "sh", "ori_t", "comp", "lvar"), r):
                ref = g[f"auto{int(auto)}_red{int(red)}/{nm}"]
                assert t.numpy().dtype == ref.dtype and np.array_equal(t.numpy(), ref), (auto, red, nm)
    # quirk Q1 is present in the fixtures (value 2 from uint8 wrap)
    assert 2 in np.unique(g["auto1_red1/comp"])
    for k in (15, 30):
        assert np.array_equal(O.dilate_mask(g["ori"][:, :, 0], k), g[f"dilate{k}"])


def test_re_edit_2d_matches_reference(golden):
    """oracle.re_edit_2d (the CPU arm's coarse edit) vs the UNMODIFIED reference's cv2 outputs: bit-exact."""
    from oracle import cases
    g = golden["coarse2d"]
    for name, (seed, ep) in cases.COARSE2D_CASES.items():
        img, m3, _, _, _ = cases.edit_case_inputs(seed, 128)
        bg, _, _, _, _ = cases.edit_case_inputs(seed + 70, 128)
        final, tmask, hole = O.re_edit_2d(img, m3, ep, bg)
        assert np.array_equal(tmask, g[name + "/tmask"]) and np.array_equal(final, g[name + "/final"]), name
        assert np.array_equal(hole, g[name + "/hole"]), name


def test_style_align_bg_matches_reference(golden):
    """style_align_share_attention_bg: 'ssa' = plain [self ; ref] attention, 'sdsa' = mask 1 - [ones ; obj] on the Q0 pairs."""
    g = golden["attention_bg"]
    T = lambda k: torch.from_numpy(g[k])
    obj = O.process_mask_before_attention(T("style_bg/obj"), 64).numpy()
    out = O.style_align(T("style_bg/q"), T("style_bg/k"), T("style_bg/v"), 8, 8 ** -0.5, None, bg=True)
    assert float((out - T("style_bg_ssa/out")).abs().max()) < 5e-5
    out = O.style_align(T("style_bg/q"), T("style_bg/k"), T("style_bg/v"), 8, 8 ** -0.5, obj, bg=True)
    assert float((out - T("style_bg_sdsa/out")).abs().max()) < 5e-5
