"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_import.py) on the seeded inputs of oracle/cases.py.

Run in the build container only:   python -m oracle.make_golden
The GPU box has no /root/reference; there the committed fixtures are the pin.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _ctrl(ref, heads, scale, method, cg, block=10):
    c = ref.attention.Attention_Modulator(start_layer=10)
    c.heads, c.scale, c.upcast_attention, c.upcast_softmax = heads, scale, False, False
    c.num_att_layers = 32
    c.cur_att_layer = 2 * block
    c.layer_idx = list(range(10, 16))
    c.method, c.context_guidance = method, cg
    return c


def gen_attention(ref):
    out = {}
    for name in cases.ATTN_CASES:
        i = cases.attn_case_inputs(name)
        c = _ctrl(ref, i["heads"], i["scale"], i["method"], i["cg"])
        c.fg_retain_mask = i["tgt"].clone()
        c.fg_retain_mask_st2 = i["tgt"].clone()
        c.fg_ref_mask = i["src"].clone()
        fn = c.Temporal_contextal_attention if i["kind"] == "edit" else c.Temporal_contextal_attention_bg
        if i["kind"] == "bg":
            c.fg_retain_mask = i["src"].clone()
        o = fn(i["q"].clone(), i["k"].clone(), i["v"].clone(), False, "up")
        # inputs are regenerated from the seed; the checksum detects RNG drift between torch versions
        out[name + "/qkv_checksum"] = np.array([float(i[t].double().sum()) for t in "qkv"] +
                                               [float(i[t].double().abs().sum()) for t in "qkv"])
        out[name + "/src"], out[name + "/tgt"] = i["src"].numpy(), i["tgt"].numpy()
        out[name + "/out"] = o.numpy()
        # down-sampled masks exactly as the reference sees them (integer work: bit-exact pin)
        out[name + "/src_ds"] = c.process_mask_before_attention(i["src"].clone(), i["S"])[0].flatten().numpy()
        out[name + "/tgt_ds"] = c.process_mask_before_attention(i["tgt"].clone(), i["S"])[0].flatten().numpy()
    # plain early-exit (block not in layer_idx) and cross-attn local modulation
    q, k, v = cases.qkv(4, 64, 64, 31)
    c = _ctrl(ref, 8, 8 ** -0.5, "tca", 0.5, block=8)
    out["plain/q"], out["plain/k"], out["plain/v"] = q.numpy(), k.numpy(), v.numpy()
    out["plain/out"] = c.Temporal_contextal_attention(q.clone(), k.clone(), v.clone(), False, "up").numpy()
    q, k, v = cases.qkv(4, 64, 64, 32, sk=77)
    region = torch.from_numpy(cases.blob_mask(64, 33))
    c = _ctrl(ref, 8, 8 ** -0.5, "tca", 0.5)
    c.local_edit_region = region.clone()
    out["cross/q"], out["cross/k"], out["cross/v"], out["cross/region"] = q.numpy(), k.numpy(), v.numpy(), region.numpy()
    out["cross/out"] = c.modulate_local_cross_attn(q.clone(), k.clone(), v.clone(), True, "up").numpy()
    # compose, N = 2 sources: streams [u_e, r1, r2, c_e]
    q, k, v = cases.qkv(4, 64, 64, 34)
    srcs = torch.from_numpy(np.stack([cases.blob_mask(64, 35), 1 - cases.blob_mask(64, 36)]))
    t0 = cases.blob_mask(64, 37)
    tgts = torch.from_numpy(np.stack([t0, 1 - t0]))
    for method, cg in (("tca", 0.3), ("mmsa", None)):
        c = _ctrl(ref, 8, 8 ** -0.5, method, cg)
        c.src_masks, c.tgt_masks = srcs.clone(), tgts.clone()
        out[f"compose_{method}/out"] = c.Temporal_contextal_attention_compose(q.clone(), k.clone(), v.clone(), False, "up").numpy()
    out["compose/q"], out["compose/k"], out["compose/v"] = q.numpy(), k.numpy(), v.numpy()
    out["compose/srcs"], out["compose/tgts"] = srcs.numpy(), tgts.numpy()
    # compose cross-attn: q [4,S,C], k/v [3 + prompt_length(2), 77, C]
    q, k, v = cases.qkv(4, 64, 64, 38, sk=77, kstreams=5)
    c = _ctrl(ref, 8, 8 ** -0.5, "tca", 0.3)
    c.tgt_masks, c.prompt_length = tgts.clone(), 2
    out["cross_compose/q"], out["cross_compose/k"], out["cross_compose/v"] = q.numpy(), k.numpy(), v.numpy()
    out["cross_compose/out"] = c.modulate_local_cross_attn_compose(q.clone(), k.clone(), v.clone(), True, "up").numpy()
    # style-align ablations (ssa: no mask; sdsa: fg_ref_mask on the ref half, Q0 tiling)
    q, k, v = cases.qkv(4, 64, 64, 39)
    src = torch.from_numpy(cases.blob_mask(64, 40))
    for method in ("ssa", "sdsa"):
        c = _ctrl(ref, 8, 8 ** -0.5, method, None)
        c.fg_ref_mask = src.clone()
        out[f"style_{method}/out"] = c.style_align_share_attention(q.clone(), k.clone(), v.clone(), False, "up").numpy()
    out["style/q"], out["style/k"], out["style/v"], out["style/src"] = q.numpy(), k.numpy(), v.numpy(), src.numpy()
    np.savez_compressed(os.path.join(OUT, "attention.npz"), **out)


def gen_style_bg(ref):
    """style_align_share_attention_bg of the UNMODIFIED reference (attention.py:1193-1238; 'sdsa' uses
    prepare_sdsa_mask_for_bggen :926-939) -- dead code in the reference's dispatch, pinned here as a method."""
    out = {}
    q, k, v = cases.qkv(4, 64, 64, 41)
    obj = torch.from_numpy(cases.blob_mask(64, 42))
    for method in ("ssa", "sdsa"):
        c = _ctrl(ref, 8, 8 ** -0.5, method, None)
        c.fg_retain_mask = obj.clone()
        out[f"style_bg_{method}/out"] = c.style_align_share_attention_bg(q.clone(), k.clone(), v.clone(), False, "up").numpy()
    out["style_bg/q"], out["style_bg/k"], out["style_bg/v"], out["style_bg/obj"] = q.numpy(), k.numpy(), v.numpy(), obj.numpy()
    np.savez_compressed(os.path.join(OUT, "attention_bg.npz"), **out)


def gen_steps(ref, parts):
    pipe, _ = ref_import.make_reference_pipeline(ref, parts)
    out = {}
    for n_steps, eta, seed in ((50, 1.0, 1), (50, 0.0, 2), (10, 1.0, 3), (10, 0.5, 4)):
        pipe.scheduler.set_timesteps(n_steps)
        eps4, x, noise, cfg_mask, var_mask = cases.step_case_inputs(seed)
        ref.model.randn_tensor = lambda shape, generator=None, device=None, dtype=None: noise.clone()
        for t in (int(pipe.scheduler.timesteps[0]), int(pipe.scheduler.timesteps[len(pipe.scheduler.timesteps) // 2]),
                  int(pipe.scheduler.timesteps[-1])):
            tt = torch.tensor(t)
            eu, ec = eps4.chunk(2)
            eps = eu + 7.5 * (ec - eu) * cfg_mask                                # model.py:610-611
            xp, x0 = pipe.ctrl_step(eps, tt, x.clone(), var_mask.clone(), eta=eta)
            xn, x0i = pipe.inv_step(eps4[:2], tt, x.clone())
            key = f"n{n_steps}_eta{eta}_t{t}"
            out[key + "/cfg"], out[key + "/x_prev"], out[key + "/pred_x0"] = eps.numpy(), xp.numpy(), x0.numpy()
            out[key + "/x_next"], out[key + "/inv_x0"] = xn.numpy(), x0i.numpy()
        out[f"seed{seed}/eps4"], out[f"seed{seed}/x"], out[f"seed{seed}/noise"] = eps4.numpy(), x.numpy(), noise.numpy()
        out[f"seed{seed}/cfg_mask"], out[f"seed{seed}/var_mask"] = cfg_mask.numpy(), var_mask.numpy()
    out["linear_param"] = np.array([[pipe.linear_param(i, 35, 50, 50, 0.0) for i in range(35, 51)],
                                    [pipe.linear_param(i, 0, 10, 16, 0.5) for i in range(0, 16)]], np.float64)
    np.savez_compressed(os.path.join(OUT, "steps.npz"), **out)


def gen_warp(ref):
    out = {}
    if ref.geo_utils is None:
        raise RuntimeError("geo_utils did not import")
    for seed in (1, 2, 3):
        src, bg, mask, M = cases.warp_case_inputs(seed)
        H, W = src.shape[-2:]
        theta = ref.geo_utils.param2theta(M, W, H)
        th = torch.tensor(theta, dtype=torch.float32)
        wb = ref.geo_utils.wrapAffine_tensor(torch.from_numpy(src), th, (W, H), mode="bilinear")
        wn = ref.geo_utils.wrapAffine_tensor(torch.from_numpy(mask.astype(np.float32)), th, (W, H), mode="nearest")
        out[f"s{seed}/src"], out[f"s{seed}/bg"], out[f"s{seed}/mask"], out[f"s{seed}/M"] = src, bg, mask, M
        out[f"s{seed}/theta"] = theta
        out[f"s{seed}/warp_bilinear"] = wb.numpy()
        out[f"s{seed}/warp_nearest"] = wn.numpy()
        wm = wn.reshape(H, W) != 0
        out[f"s{seed}/blend"] = torch.where(wm[None, None], wb.reshape(src.shape), torch.from_numpy(bg)).numpy()
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **out)


def gen_masks(ref, parts):
    pipe, _ = ref_import.make_reference_pipeline(ref, parts)
    out = {}
    res, lat = 128, 16
    ori = cases.blob_mask(res, 51, (0.08, 0.15))
    shifted = np.roll(ori, (12, -17), (0, 1)) * 255             # re_edit_2d returns 0/255
    draw = cases.blob_mask(res, 52, (0.1, 0.2))
    cons = cases.blob_mask(res, 53, (0.05, 0.1)) * 255
    ori3 = np.repeat(ori[:, :, None], 3, 2)                     # read_and_resize_mask: 3-channel 0/1
    init = torch.zeros(1, 4, lat, lat)
    out["ori"], out["shifted"], out["draw"], out["cons"] = ori3, shifted, draw, cons
    for auto in (False, True):
        for red in (False, True):
            r = pipe.prepare_various_mask(shifted.copy(), ori3.copy(), draw.copy(), res, res, init, verbose=True,
                                          use_auto_draw=auto, cons_area=cons.copy(), reduce_inp_artifacts=red)
            for nm, t in zip(("fg", "sh", "ori_t", "comp", "lvar"), r):
                out[f"auto{int(auto)}_red{int(red)}/{nm}"] = t.numpy()
    for k in (15, 30):
        out[f"dilate{k}"] = pipe.dilate_mask(ori.copy(), k)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), **out)


def gen_pipeline(ref):
    """Whole edits through the UNMODIFIED reference pipeline (DDIM_inversion_func -> Details_Preserving_regeneration,
    i.e. FreeFine_generation model.py:1012-1049 without its GIF writer) on the tiny stand-in network, CPU fp32."""
    from freefine_b200.standin import build_standin
    out = {}
    for name, c in cases.PIPE_CASES.items():
        parts = build_standin("tiny")
        pipe, controller = ref_import.make_reference_pipeline(ref, parts)
        img, ori_mask3, edit_param, draw, cons = cases.edit_case_inputs(c["seed"], c["res"])
        coarse, tgt_mask, _ = ref.vis_utils.re_edit_2d(img, ori_mask3, edit_param, img)
        if c.get("driver_like"):          # freefine_batch_infer_2d.py:196,229
            draw, cons = np.ones_like(ori_mask3[:, :, 0]), tgt_mask
        counter = {"k": 0}

        def fake_randn(shape, generator=None, device=None, dtype=None, _c=c, _n=counter):
            t = cases.step_noise(_c["seed"], _n["k"], shape)
            _n["k"] += 1
            return t

        ref.model.randn_tensor = fake_randn
        torch.manual_seed(c["seed"])
        ori_mask = pipe.mask_reduce_dim(ori_mask3)
        _, inv = pipe.DDIM_inversion_func(img=coarse, mask=tgt_mask, prompt="", num_step=c["num_step"],
                                          start_step=c["start_step"], ref_img=img, verbose=True)
        edit_img, ref_img, inter = pipe.Details_Preserving_regeneration(
            coarse, inv, c["prompt"], tgt_mask, ori_mask, draw, num_steps=c["num_step"], start_step=c["start_step"],
            end_step=c["end_step"], guidance_scale=c["gs"], eta=c["eta"], share_attn=True, method_type=c["method"],
            verbose=True, local_text_edit=True, local_perturbation=True, return_intermediates=True,
            cons_area=cons, use_auto_draw=c["use_auto_draw"], end_scale=c["end_scale"],
            reduce_inp_artifacts=c["reduce_inp_artifacts"])
        out[name + "/img"], out[name + "/ori_mask"], out[name + "/coarse"], out[name + "/tgt_mask"] = img, ori_mask3, coarse, tgt_mask
        out[name + "/draw"], out[name + "/cons"] = draw, cons
        out[name + "/inverted"] = torch.stack(inv).numpy()
        out[name + "/latents"] = torch.stack(inter).numpy()
        out[name + "/edit_img"], out[name + "/ref_img"] = edit_img, ref_img
        out[name + "/n_noise"] = np.array(counter["k"])
        out[name + "/params_json"] = np.array(json.dumps(c))      # read by freefine_b200/selfcheck.py (no oracle import there)
    # background generation / object removal (register_attention_control_4bggen, model.py:1088-1118 without the GIF)
    for name, c in cases.BG_CASES.items():
        parts = build_standin("tiny")
        pipe, controller = ref_import.make_reference_pipeline(ref, parts, flavour="bggen")
        img, ori_mask3, _, _, _ = cases.edit_case_inputs(c["seed"], c["res"])
        counter = {"k": 0}

        def fake_randn(shape, generator=None, device=None, dtype=None, _c=c, _n=counter):
            t = cases.step_noise(_c["seed"], _n["k"], shape)
            _n["k"] += 1
            return t

        ref.model.randn_tensor = fake_randn
        torch.manual_seed(c["seed"])
        ori_mask = pipe.mask_reduce_dim(ori_mask3)
        _, inv = pipe.DDIM_inversion_func(img=img, mask=ori_mask, prompt="", num_step=c["num_step"], start_step=c["start_step"],
                                          ref_img=None, verbose=True)
        edit_img, inter = pipe.Details_Preserving_regeneration_background(
            img, inv, c["prompt"], ori_mask, num_steps=c["num_step"], start_step=c["start_step"], end_step=c["end_step"],
            guidance_scale=c["gs"], eta=c["eta"], verbose=True, end_scale=c["end_scale"], return_intermediates=True,
            method_type=c["method"])
        out[name + "/img"], out[name + "/ori_mask"] = img, ori_mask3
        out[name + "/inverted"] = torch.stack(inv).numpy()
        out[name + "/latents"] = torch.stack([x.reshape(4, *x.shape[-2:]) for x in inter]).numpy()
        out[name + "/edit_img"] = edit_img
    # cross-image composition / appearance transfer (register_attention_control_compose; entered at the inner functions:
    # the public entry point raises TypeError as published)
    for name, c in cases.COMPOSE_CASES.items():
        parts = build_standin("tiny")
        pipe, controller = ref_import.make_reference_pipeline(ref, parts, flavour="compose")
        ci = cases.compose_case_inputs(c["seed"], c["res"])
        counter = {"k": 0}

        def fake_randn(shape, generator=None, device=None, dtype=None, _c=c, _n=counter):
            t = cases.step_noise(_c["seed"], _n["k"], shape)
            _n["k"] += 1
            return t

        ref.model.randn_tensor = fake_randn
        torch.manual_seed(c["seed"])
        inv = pipe.DDIM_inversion_func_compose(img=ci["coarse"], compose_imgs=ci["imgs"], prompt="", num_step=c["num_step"],
                                               start_step=c["start_step"], verbose=True)
        image, inter = pipe.Details_Preserving_regeneration_compose(
            ci["coarse"], inv, list(c["prompt"]), [m.copy() for m in ci["ori_masks"]], [m.copy() for m in ci["tgt_masks"]], None,
            num_steps=c["num_step"], start_step=c["start_step"], end_step=c["end_step"], eta=c["eta"], guidance_scale=c["gs"],
            dil_completion=c["dil_completion"], appearance_transfer=c["appearance_transfer"], method_type=c["method"],
            verbose=True, return_intermediates=True, end_scale=c["end_scale"])
        out[name + "/inverted"] = torch.stack(inv).numpy()
        out[name + "/latents"] = torch.stack([x.reshape(4, *x.shape[-2:]) for x in inter[1:]]).numpy()
        out[name + "/edit_img"] = image
        out[name + "/tgt_masks"] = controller.tgt_masks.numpy() if controller.tgt_masks is not None else np.zeros(1)
    np.savez_compressed(os.path.join(OUT, "pipeline.npz"), **out)


def gen_config1(ref):
    """BASELINE.json configs[0] / SURVEY.md 8d config 1: the reference's own example image (Examples/Editing/2D/bear,
    read like vis_utils.py:349-360), moved by dx = +60 px, 10-step inversion + 10-step TCA sampling, tiny stand-in UNet
    on the CPU.  The two PNGs are copied beside the fixture so that the GPU test reads the same files."""
    import shutil
    from freefine_b200.standin import build_standin
    c = cases.CONFIG1
    src_dir = os.path.join(ref_import.REF_ROOT, "Examples", "Editing", "2D", c["example"])
    for f in ("source.png", "source_mask.png"):
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(OUT, f"config1_{c['example']}_{f}"))
    img = ref.vis_utils.read_and_resize_img(os.path.join(src_dir, "source.png"))
    ori_mask3 = ref.vis_utils.read_and_resize_mask(os.path.join(src_dir, "source_mask.png"))
    coarse, tgt_mask, _ = ref.vis_utils.re_edit_2d(img, ori_mask3, c["edit_param"], img)
    parts = build_standin("tiny")
    pipe, controller = ref_import.make_reference_pipeline(ref, parts)
    counter = {"k": 0}

    def fake_randn(shape, generator=None, device=None, dtype=None):
        t = cases.step_noise(c["seed"], counter["k"], shape)
        counter["k"] += 1
        return t

    ref.model.randn_tensor = fake_randn
    torch.manual_seed(c["seed"])
    ori_mask = pipe.mask_reduce_dim(ori_mask3)
    _, inv = pipe.DDIM_inversion_func(img=coarse, mask=tgt_mask, prompt="", num_step=c["num_step"],
                                      start_step=c["start_step"], ref_img=img, verbose=True)
    edit_img, ref_img, inter = pipe.Details_Preserving_regeneration(
        coarse, inv, c["prompt"], tgt_mask, ori_mask, np.ones_like(ori_mask), num_steps=c["num_step"],
        start_step=c["start_step"], end_step=c["end_step"], guidance_scale=c["gs"], eta=c["eta"], share_attn=True,
        method_type=c["method"], verbose=True, local_text_edit=True, local_perturbation=True, return_intermediates=True,
        cons_area=tgt_mask, use_auto_draw=c["use_auto_draw"], end_scale=c["end_scale"],
        reduce_inp_artifacts=c["reduce_inp_artifacts"])
    out = {"img_checksum": np.array([int(img.astype(np.int64).sum()), int(ori_mask.astype(np.int64).sum())]),
           "tgt_mask_bits": np.packbits(tgt_mask != 0), "coarse_checksum": np.array(int(coarse.astype(np.int64).sum())),
           "inverted_last": inv[-1].numpy(), "latents_last": inter[-1].numpy(),
           "latents_mid": inter[len(inter) // 2].numpy(), "edit_img_small": edit_img[::4, ::4].copy(),
           "n_noise": np.array(counter["k"]), "n_latents": np.array(len(inter)), "params_json": np.array(json.dumps(c))}
    np.savez_compressed(os.path.join(OUT, "config1.npz"), **out)


def gen_coarse2d(ref):
    """re_edit_2d of the UNMODIFIED reference (cv2.warpAffine) incl. rotation + anisotropic scale."""
    out = {}
    for name, (seed, ep) in cases.COARSE2D_CASES.items():
        img, m3, _, _, _ = cases.edit_case_inputs(seed, 128)
        bg, _, _, _, _ = cases.edit_case_inputs(seed + 70, 128)
        final, tmask, hole = ref.vis_utils.re_edit_2d(img, m3, ep, bg)
        out[name + "/final"], out[name + "/tmask"], out[name + "/hole"] = final, tmask, hole
    np.savez_compressed(os.path.join(OUT, "coarse2d.npz"), **out)


def gen_coarse3d(ref):
    """re_edit_3d of the UNMODIFIED reference (cv2.warpAffine) on seeded inputs."""
    out = {}
    for name, (seed, ep) in cases.COARSE3D_CASES.items():
        src, m3, bg, ori, om = cases.coarse3d_case_inputs(seed)
        final, tmask, hole = ref.vis_utils.re_edit_3d(src, m3, ep, bg, ori, om)
        out[name + "/final"], out[name + "/tmask"], out[name + "/hole"] = final, tmask, hole
    np.savez_compressed(os.path.join(OUT, "coarse3d.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.load()
    from freefine_b200.standin import build_standin
    parts = build_standin("tiny")
    torch.set_grad_enabled(False)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:                                    # e.g. `python -m oracle.make_golden config1 pipeline`
        for name in only:
            fn = globals()["gen_" + name]
            fn(ref, parts) if name in ("steps", "masks") else fn(ref)
        return
    gen_attention(ref)
    gen_style_bg(ref)
    gen_steps(ref, parts)
    gen_warp(ref)
    gen_masks(ref, parts)
    gen_pipeline(ref)
    gen_coarse3d(ref)
    gen_coarse2d(ref)
    gen_config1(ref)
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
