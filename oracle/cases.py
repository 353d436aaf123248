"""TEST INFRASTRUCTURE ONLY -- seeded input generators shared by oracle/make_golden.py and tests/.

Every case is fully determined by its name (seeds are explicit) so that the fixture generator (run in the build
container against the unmodified reference) and the parity tests (run anywhere) see identical inputs.  Inputs
are ALSO stored inside the fixtures, so a change in a library RNG cannot silently unpin the goldens.
"""
from __future__ import annotations

import numpy as np
import torch


def blob_mask(res: int, seed: int, frac=(0.15, 0.3), dtype=np.uint8) -> np.ndarray:
    """Filled ellipse, uint8 0/1 (SURVEY.md 8d synthetic source mask)."""
    rng = np.random.default_rng(seed)
    cy, cx = rng.uniform(0.3 * res, 0.7 * res, 2)
    ay, ax = rng.uniform(frac[0] * res, frac[1] * res, 2)
    yy, xx = np.mgrid[0:res, 0:res]
    return ((((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2) <= 1.0).astype(dtype)


def qkv(streams: int, S: int, C: int, seed: int, sk: int | None = None, logit_scale: float = 3.0, kstreams=None):
    """q,k,v [B,S,C] fp32 with |logit| of a few units (random-init nets give near-uniform softmax: too easy).
    Values are rounded to the bf16 grid (still stored as fp32) so that the SAME numbers can be fed to the fp32
    reference / oracle and, exactly, to the bf16 tensor-core kernel."""
    g = torch.Generator().manual_seed(seed)
    sk = S if sk is None else sk
    kstreams = streams if kstreams is None else kstreams
    q = torch.randn(streams, S, C, generator=g) * logit_scale
    k = torch.randn(kstreams, sk, C, generator=g)
    v = torch.randn(kstreams, sk, C, generator=g)
    return tuple(t.bfloat16().float() for t in (q, k, v))


# name -> dict(kind=..., params)
ATTN_CASES = {
    # (kind, method, heads, d, res(full mask), S, cg, src_mode)
    "tca_h8_s256":        dict(kind="edit", method="tca",  heads=8, d=8,  res=128, S=256, cg=0.6, src="blob", seed=11),
    "mmsa_h8_s256":       dict(kind="edit", method="mmsa", heads=8, d=8,  res=128, S=256, cg=None, src="blob", seed=12),
    "tca_h2_s64_d40":     dict(kind="edit", method="tca",  heads=2, d=40, res=64,  S=64,  cg=0.25, src="blob", seed=13),
    "tca_h8_s64_empty":   dict(kind="edit", method="tca",  heads=8, d=8,  res=64,  S=64,  cg=1.0, src="empty", seed=14),
    "tca_h8_s64_full":    dict(kind="edit", method="tca",  heads=8, d=8,  res=64,  S=64,  cg=0.5, src="full", seed=15),
    "tca_h8_s64_onekey":  dict(kind="edit", method="tca",  heads=8, d=8,  res=64,  S=64,  cg=0.9, src="one", seed=16),
    "bg_tca_h8_s256":     dict(kind="bg",   method="tca",  heads=8, d=8,  res=128, S=256, cg=0.7, src="blob", seed=17),
    "bg_mmsa_h8_s64":     dict(kind="bg",   method="mmsa", heads=8, d=8,  res=64,  S=64,  cg=None, src="blob", seed=18),
    "tca_h8_s256_res256": dict(kind="edit", method="tca", heads=8, d=8,  res=256, S=256, cg=0.4, src="blob", seed=19),
    # SD1.5 head dims (40: up3/down0, 80: up2/down1, 160: up1/mid) at small token counts
    "tca_h8_s256_d40":    dict(kind="edit", method="tca",  heads=8, d=40,  res=128, S=256, cg=0.7, src="blob", seed=20),
    "tca_h8_s64_d80":     dict(kind="edit", method="tca",  heads=8, d=80,  res=64,  S=64,  cg=0.35, src="blob", seed=21),
    "mmsa_h8_s64_d160":   dict(kind="edit", method="mmsa", heads=8, d=160, res=64,  S=64,  cg=None, src="blob", seed=22),
    "bg_tca_h8_s64_d160": dict(kind="bg",   method="tca",  heads=8, d=160, res=64,  S=64,  cg=0.5, src="blob", seed=23),
}


def attn_case_inputs(name: str):
    c = ATTN_CASES[name]
    C = c["heads"] * c["d"]
    q, k, v = qkv(4, c["S"], C, c["seed"])
    res = c["res"]
    tgt = blob_mask(res, c["seed"] + 100)
    if c["src"] == "blob":
        src = blob_mask(res, c["seed"] + 200)
    elif c["src"] == "empty":
        src = np.zeros((res, res), np.uint8)
    elif c["src"] == "full":
        src = np.ones((res, res), np.uint8)
    else:
        src = np.zeros((res, res), np.uint8)
        src[res // 2, res // 2] = 1          # survives nearest down-sampling only if on the sampling lattice
        src[0, 0] = 1
    out = dict(c)
    out.update(q=q, k=k, v=v, src=torch.from_numpy(src), tgt=torch.from_numpy(tgt), scale=c["d"] ** -0.5)
    return out


def step_case_inputs(seed: int, h=16, w=16):
    g = torch.Generator().manual_seed(seed)
    eps4 = torch.randn(4, 4, h, w, generator=g)
    x = torch.randn(2, 4, h, w, generator=g)
    noise = torch.randn(2, 4, h, w, generator=g)
    rng = np.random.default_rng(seed)
    cfg_mask = torch.from_numpy(rng.integers(0, 3, (h, w)).astype(np.uint8))     # {0,1,2}: quirk Q1
    var_mask = torch.from_numpy(rng.integers(0, 3, (h, w)).astype(np.uint8))
    return eps4, x, noise, cfg_mask, var_mask


def warp_case_inputs(seed: int, N=1, C=8, H=32, W=32):  # reference wrapAffine_tensor only supports N=1 (theta.unsqueeze(0))
    rng = np.random.default_rng(seed)
    src = rng.standard_normal((N, C, H, W)).astype(np.float32)
    bg = rng.standard_normal((N, C, H, W)).astype(np.float32)
    mask = blob_mask(H, seed + 5)
    ang = rng.uniform(-30, 30)
    sc = rng.uniform(0.7, 1.3)
    dx, dy = rng.uniform(-0.15 * W, 0.15 * W, 2)
    cx, cy = W / 2 - 0.5, H / 2 - 0.5
    a, b = sc * np.cos(np.deg2rad(ang)), sc * np.sin(np.deg2rad(ang))
    M = np.array([[a, b, (1 - a) * cx - b * cy + dx], [-b, a, b * cx + (1 - a) * cy + dy]], np.float64)
    return src, bg, mask, M


def edit_case_inputs(seed: int, res: int = 128):
    """One synthetic 2-D edit: smooth random image, elliptical object mask, a translation that keeps
    the object inside.  Returns (image uint8 [res,res,3], ori_mask uint8 0/1 [res,res,3], edit_param 5-tuple,
    draw_mask uint8 0/1 [res,res], cons_area uint8 0/255 [res,res])."""
    rng = np.random.default_rng(1000 + seed)
    img = rng.integers(0, 256, (res, res, 3)).astype(np.float32)
    k = max(3, res // 16) | 1                     # separable box blur: smooth, non-trivial content (no cv2 needed)
    ker = np.ones(k, np.float32) / k
    for ax in (0, 1):
        img = np.apply_along_axis(lambda v: np.convolve(np.pad(v, k // 2, mode="edge"), ker, mode="valid"), ax, img)
    img = np.clip(img, 0, 255).astype(np.uint8)
    m = blob_mask(res, 2000 + seed, (0.08, 0.16))
    ys, xs = np.where(m)
    lim = lambda lo, hi: (-(lo - 2), (res - 3) - hi)
    dxl, dxh = lim(xs.min(), xs.max())
    dyl, dyh = lim(ys.min(), ys.max())
    dx = int(np.clip(rng.integers(-res // 5, res // 5 + 1), dxl, dxh))
    dy = int(np.clip(rng.integers(-res // 5, res // 5 + 1), dyl, dyh))
    if dx == 0 and dy == 0:
        dx = int(np.clip(res // 8, dxl, dxh))
    draw = blob_mask(res, 3000 + seed, (0.1, 0.2))
    cons = blob_mask(res, 4000 + seed, (0.05, 0.1)) * 255
    return img, np.repeat(m[:, :, None], 3, 2), (dx, dy, 0, 1.0, 1.0), draw, cons


def step_noise(seed: int, k: int, shape):
    """k-th randn_tensor draw of an edit (fed to both the reference and the B200 path)."""
    return torch.randn(tuple(shape), generator=torch.Generator().manual_seed(42 + 1000 * seed + k))


PIPE_CASES = {
    # name -> kwargs of FreeFine_generation's inner calls (tiny stand-in UNet, res 128 -> 16x16 latents)
    "quirk_free":   dict(seed=1, res=128, num_step=10, start_step=2, end_step=6, eta=1.0, gs=7.5, method="tca",
                         use_auto_draw=False, reduce_inp_artifacts=False, end_scale=0.5, prompt="a photo of a thing"),
    "geobench_2d":  dict(seed=2, res=128, num_step=10, start_step=4, end_step=10, eta=1.0, gs=7.5, method="tca",
                         use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0, prompt=""),
    "mmsa":         dict(seed=3, res=128, num_step=8, start_step=2, end_step=8, eta=0.0, gs=5.0, method="mmsa",
                         use_auto_draw=False, reduce_inp_artifacts=False, end_scale=0.5, prompt="a photo of a thing"),
    # the full 50-step schedule (BASELINE.json north_star: "final latents after a full 50-step edit"), driver-like inputs
    # (cons_area = target mask, draw_mask = ones: evaluation/FreeFine/freefine_batch_infer_2d.py:196,212-230):
    # the GeoBench-2D default (15 + 15 UNet calls), the whole schedule (50 + 50, what bench.py times by default) with the
    # quirk-faithful settings, and the whole schedule quirk-free
    "sched50_ss35": dict(seed=8, res=128, num_step=50, start_step=35, end_step=50, eta=1.0, gs=7.5, method="tca",
                         use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0, prompt="", driver_like=True),
    "sched50_full": dict(seed=9, res=128, num_step=50, start_step=0, end_step=50, eta=1.0, gs=7.5, method="tca",
                         use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0, prompt="", driver_like=True),
    # the style-align ablation methods through forward_sampling (model.py:514-520: every self-attention layer attends to
    # [self ; ref] keys, SDSA with the source-object mask on the ref half)
    "ssa":  dict(seed=12, res=128, num_step=6, start_step=1, end_step=4, eta=1.0, gs=7.5, method="ssa",
                 use_auto_draw=False, reduce_inp_artifacts=False, end_scale=0.5, prompt="a photo of a thing"),
    "sdsa": dict(seed=13, res=128, num_step=6, start_step=1, end_step=4, eta=0.0, gs=7.5, method="sdsa",
                 use_auto_draw=False, reduce_inp_artifacts=False, end_scale=0.5, prompt="a photo of a thing"),
    "sched50_quirkfree": dict(seed=10, res=128, num_step=50, start_step=0, end_step=25, eta=1.0, gs=7.5, method="tca",
                              use_auto_draw=False, reduce_inp_artifacts=False, end_scale=0.5, prompt="a photo of a thing"),
}

# BASELINE.json configs[0]: object repositioning on one 512x512 Examples/Editing image, 10-step inversion + sampling
# (SURVEY.md 8d "Config 1"): Examples/Editing/2D/bear, dx = +60 px, GeoBench-2D settings otherwise.
CONFIG1 = dict(example="bear", edit_param=(60, 0, 0, 1.0, 1.0), num_step=10, start_step=0, end_step=10, eta=1.0, gs=7.5,
               method="tca", use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0, prompt="", seed=11)


BG_CASES = {
    "bg_tca":  dict(seed=4, res=128, num_step=8, start_step=1, end_step=5, eta=1.0, gs=7.5, method="tca", end_scale=0.5,
                    prompt="empty scene"),
    "bg_mmsa": dict(seed=5, res=128, num_step=6, start_step=1, end_step=6, eta=0.0, gs=7.5, method="mmsa", end_scale=0.5,
                    prompt="empty scene"),
}


COMPOSE_CASES = {
    # appearance transfer (model.py:1516-1539): refs = [appearance image, original image],
    # src masks = [app_mask, 1-ori_mask], tgt masks = [ori_mask] (+ background appended by prepare_composition_masks)
    "appearance": dict(seed=6, res=128, num_step=8, start_step=2, end_step=6, eta=1.0, gs=7.5, method="tca", end_scale=0.5,
                       appearance_transfer=True, dil_completion=False, prompt=["a photo of a thing"]),
    "compose_mmsa": dict(seed=7, res=128, num_step=6, start_step=1, end_step=6, eta=0.0, gs=5.0, method="mmsa", end_scale=0.5,
                         appearance_transfer=False, dil_completion=True, prompt=["a thing"]),
}


def compose_case_inputs(seed: int, res: int = 128):
    img, ori_mask3, _, _, _ = edit_case_inputs(seed, res)
    app_img, app_mask3, _, _, _ = edit_case_inputs(seed + 10, res)
    ori = ori_mask3[:, :, 0]
    return dict(coarse=img, imgs=[app_img, img], ori_masks=[app_mask3[:, :, 0].copy(), (1 - ori).astype(np.uint8)],
                tgt_masks=[ori.copy()])


# re_edit_3d (vis_utils.py:275-339): (seed, edit_param) -- translation only, and translation + rotation + anisotropic scale
COARSE3D_CASES = {
    "move": (31, (9, -6, 0, 1.0, 1.0)),
    "move_rot_scale": (32, (-7, 5, 17.0, 0.85, 1.2)),
}


def coarse3d_case_inputs(seed: int, res: int = 128):
    """(re-oriented object image, its mask [res,res,3] 0/1, background, original image, original mask [res,res,1] bool)."""
    src, m3, _, _, _ = edit_case_inputs(seed, res)
    ori, om3, _, _, _ = edit_case_inputs(seed + 50, res)
    bg, _, _, _, _ = edit_case_inputs(seed + 70, res)
    return src, m3, bg, ori, om3[:, :, :1].astype(bool)


# re_edit_2d (vis_utils.py:210-274) with rotation + anisotropic scale: (seed, edit_param (dx, dy, rz, sx, sy))
COARSE2D_CASES = {
    "move": (41, (11, -7, 0, 1.0, 1.0)),
    "rot_scale": (42, (-6, 9, 23.0, 1.15, 0.85)),
    "rot_only": (43, (0, 0, -31.0, 1.0, 1.0)),
}
