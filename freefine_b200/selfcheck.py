"""Whole-edit parity check of the B200 path against the committed reference goldens (tests/golden/pipeline.npz,
tests/golden/config1.npz -- final latents produced by the UNMODIFIED reference on the CPU, oracle/make_golden.py).

Used by the `-m gpu` tests and by bench.py's `parity` block, so that the number reported next to a throughput is
measured in the same process, on the same UNet path (bf16 channels-last fast path or fp32) that was timed.  Everything
an edit needs -- inputs, parameters, the seeds of the noise draws -- is read from the fixture; nothing under oracle/ is
imported (the oracle is test infrastructure, the fixtures are data).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def step_noise(seed: int, k: int, shape, device=None):
    """k-th randn_tensor draw of golden edit `seed` (the generator of oracle/cases.py::step_noise, which fed the
    reference when the fixture was made)."""
    t = torch.randn(tuple(shape), generator=torch.Generator().manual_seed(42 + 1000 * seed + k))
    return t if device is None else t.to(device)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm())


def build_pipeline(device, dtype=torch.float32, preset="tiny"):
    """Stand-in parts + controller + registered attention plugin, as evaluation/FreeFine/freefine_batch_infer_2d.py:149-155."""
    from .pipeline import Attention_Modulator, FreeFinePipeline, register_attention_control
    from .standin import build_standin
    parts = build_standin(preset, device=device, dtype=dtype)
    controller = Attention_Modulator(start_layer=10)
    pipe = FreeFinePipeline.from_parts(parts, controller, device=device)
    register_attention_control(pipe, controller)
    pipe.modify_unet_forward()
    return pipe, controller


class _Noise:
    def __init__(self, seed):
        self.seed, self.k = seed, 0

    def __call__(self, shape, generator=None, device=None, dtype=None):
        t = step_noise(self.seed, self.k, shape, device)
        self.k += 1
        return t


def run_golden_edit(pipe, name: str, golden=None):
    """Runs fixture edit `name` (DDIM_inversion_func -> Details_Preserving_regeneration, i.e. FreeFine_generation
    model.py:1012-1049) through `pipe` and returns a dict of relative-L2 errors against the reference's latents."""
    from . import pipeline as P
    g = golden if golden is not None else np.load(os.path.join(GOLDEN_DIR, "pipeline.npz"))
    c = json.loads(str(g[name + "/params_json"]))
    noise = _Noise(c["seed"])
    old, P.randn_tensor = P.randn_tensor, noise
    try:
        img, coarse, tgt_mask = g[name + "/img"], g[name + "/coarse"], g[name + "/tgt_mask"]
        ori_mask = pipe.mask_reduce_dim(g[name + "/ori_mask"])
        _, inv = pipe.DDIM_inversion_func(img=coarse, mask=tgt_mask, prompt="", num_step=c["num_step"],
                                          start_step=c["start_step"], ref_img=img, verbose=True)
        edit_img, ref_img, inter = pipe.Details_Preserving_regeneration(
            coarse, inv, c["prompt"], tgt_mask, ori_mask, g[name + "/draw"], num_steps=c["num_step"],
            start_step=c["start_step"], end_step=c["end_step"], guidance_scale=c["gs"], eta=c["eta"], share_attn=True,
            method_type=c["method"], verbose=True, local_text_edit=True, local_perturbation=True,
            return_intermediates=True, cons_area=g[name + "/cons"], use_auto_draw=c["use_auto_draw"],
            end_scale=c["end_scale"], reduce_inp_artifacts=c["reduce_inp_artifacts"])
    finally:
        P.randn_tensor = old
    inv_ref = torch.from_numpy(g[name + "/inverted"])
    lat_ref = torch.from_numpy(g[name + "/latents"])
    errs = [rel_l2(a.float().cpu(), b) for a, b in zip(inter, lat_ref)]
    return dict(name=name, n_inverted=len(inv), n_inverted_ref=int(inv_ref.shape[0]), n_latents=len(inter),
                n_latents_ref=int(lat_ref.shape[0]), n_noise=noise.k, n_noise_ref=int(g[name + "/n_noise"]),
                inverted_rel_l2=rel_l2(inv[-1].float().cpu(), inv_ref[-1]), final_rel_l2=errs[-1], max_rel_l2=max(errs),
                img_max_abs=int(np.abs(edit_img.astype(np.int32) - g[name + "/edit_img"].astype(np.int32)).max()),
                edit_img=edit_img)


def run_config1(pipe, golden=None):
    """BASELINE.json configs[0]: the reference's own example image (tests/golden/config1_bear_source*.png, copies of
    Examples/Editing/2D/bear), read like vis_utils.py:349-360, moved by dx = +60 px with re_edit_2d on the warp kernel,
    10-step inversion + 10-step TCA sampling; compared with the reference's latents (tests/golden/config1.npz)."""
    from . import pipeline as P
    from .coarse_edit import re_edit_2d
    from .geobench import read_and_resize_img, read_and_resize_mask
    g = golden if golden is not None else np.load(os.path.join(GOLDEN_DIR, "config1.npz"))
    c = json.loads(str(g["params_json"]))
    img = read_and_resize_img(os.path.join(GOLDEN_DIR, f"config1_{c['example']}_source.png"))
    ori_mask = read_and_resize_mask(os.path.join(GOLDEN_DIR, f"config1_{c['example']}_source_mask.png"))
    inputs_ok = [int(img.astype(np.int64).sum()), int(ori_mask.astype(np.int64).sum())] == [int(x) for x in g["img_checksum"]]
    coarse, tgt_mask, _ = re_edit_2d(img, ori_mask, tuple(c["edit_param"]), img)
    mask_exact = bool(np.array_equal(np.packbits(tgt_mask != 0), g["tgt_mask_bits"]))
    noise = _Noise(c["seed"])
    old, P.randn_tensor = P.randn_tensor, noise
    try:
        _, inv = pipe.DDIM_inversion_func(img=coarse, mask=tgt_mask, prompt="", num_step=c["num_step"],
                                          start_step=c["start_step"], ref_img=img, verbose=True)
        edit_img, ref_img, inter = pipe.Details_Preserving_regeneration(
            coarse, inv, c["prompt"], tgt_mask, ori_mask, np.ones_like(ori_mask), num_steps=c["num_step"],
            start_step=c["start_step"], end_step=c["end_step"], guidance_scale=c["gs"], eta=c["eta"], share_attn=True,
            method_type=c["method"], verbose=True, local_text_edit=True, local_perturbation=True,
            return_intermediates=True, cons_area=tgt_mask, use_auto_draw=c["use_auto_draw"], end_scale=c["end_scale"],
            reduce_inp_artifacts=c["reduce_inp_artifacts"])
    finally:
        P.randn_tensor = old
    return dict(name="config1_" + c["example"], inputs_ok=inputs_ok, coarse_mask_bit_exact=mask_exact,
                n_latents=len(inter), n_latents_ref=int(g["n_latents"]), n_noise=noise.k, n_noise_ref=int(g["n_noise"]),
                inverted_rel_l2=rel_l2(inv[-1].float().cpu(), torch.from_numpy(g["inverted_last"])),
                mid_rel_l2=rel_l2(inter[len(inter) // 2].float().cpu(), torch.from_numpy(g["latents_mid"])),
                final_rel_l2=rel_l2(inter[-1].float().cpu(), torch.from_numpy(g["latents_last"])),
                coarse_checksum_delta=int(coarse.astype(np.int64).sum()) - int(g["coarse_checksum"]))
