"""GeoBench-2D wire formats and the sharded batch-inference driver -- the caller of the hot path.

Mirrors evaluation/FreeFine/freefine_batch_infer_2d.py of the reference:
* `flatten_cases`   = CustomDataset.__init__ (:88-113): `annotations_2d.json` is
      {da_n: {"instances": {ins_id: {edit_ins: {ori_img_path, ori_mask_path, edit_param[9], ...}}}}}
  and flattens, in file order, into items {da_n, ins_id, edit_ins, **input_pack}; an item whose output PNG
  `<gen_dir>/<da_n>/<ins_id>/<edit_ins>.png` already exists is an *existing result* (resume) and gets `gen_img_path`.
* `shard`           = DistributedSampler(shuffle=False, drop_last=False) (:167), via freefine_b200.dist.shard_indices.
* `merge_results`   = the rank-0 merge (:245-262): existing results first, then every rank's list in rank order, into
      {da_n: {"instances": {ins_id: {edit_ins: item}}}} -> `generated_results_freefine_2d.json` (indent 4, utf-8).
  The sampler pads the tail ranks with repeated samples; a repeated item simply overwrites its own key, exactly as in
  the reference.
* `run`             = main (:141-262) with E edits per stream batch instead of batch_size 1: reads images / masks like
  read_and_resize_img / read_and_resize_mask (src/utils/vis_utils.py:349-360), hands (image, mask, edit_param,
  inpainted background) to FreeFinePipeline.FreeFine_generation_batch -- which performs re_edit_2d on the GPU -- with the
  reference's GeoBench-2D settings (:212-230), writes PNGs like save_img (:117-134) and gathers the metadata with
  all_gather_object (:243).  Reference quirk: main() calls re_edit_2d with `ori_mask` before assigning it (:190 vs :192,
  a NameError as published); here the mask is read first.
"""
from __future__ import annotations

import json
import os
import os.path as osp

import numpy as np

from . import dist as ffdist

GEN_SUBDIR = osp.join("Geo-Bench-2D", "Gen_results_FreeFine_2d")
INP_SUBDIR = osp.join("Geo-Bench-2D", "inp_img_blended")
RESULT_JSON = "generated_results_freefine_2d.json"
RESULT_JSON_3D = "generated_results_freefine_depth.json"
GEN_SUBDIR_3D = osp.join("Geo-Bench-3D", "Gen_results_FreeFine_depth")
COARSE_SUBDIR_3D = "coarse3d_depth_anything"
# freefine_batch_infer_3d_depth.py:144-162 (config 5's caller): the coarse input comes from the depth pre-step, the user's
# draw mask replaces the automatic completion region, the constraint area is the target mask
GEOBENCH_3D_SETTINGS = dict(guidance_scale=7.5, eta=1.0, end_scale=0.0, end_step=50, num_step=50, start_step=15, seed=42,
                            use_auto_draw=False, reduce_inp_artifacts=True)
# freefine_batch_infer_bggen_2d.py:165-179 (the background every 2-D edit is blended onto)
BGGEN_2D_SETTINGS = dict(guidance_text="empty scene", guidance_scale=7.5, eta=1.0, end_scale=0.5, end_step=35, num_step=50,
                         start_step=1, use_auto_draw=False, reduce_inp_artifacts=True)
GEOBENCH_2D_SETTINGS = dict(guidance_scale=7.5, eta=1.0, end_scale=0.0, end_step=50, num_step=50, start_step=35, seed=42,
                            use_auto_draw=True, reduce_inp_artifacts=True)


def expected_path(gen_dir: str, da_n, ins_id, edit_ins, make_dirs: bool = True) -> str:
    d = osp.join(gen_dir, str(da_n), str(ins_id))
    if make_dirs:
        os.makedirs(d, exist_ok=True)
    return osp.join(d, f"{edit_ins}.png")


def flatten_cases(data: dict, gen_dir: str, check_exist: bool = True, make_dirs: bool = True):
    """-> (cases, existing_results), both lists of item dicts in annotation order."""
    cases, existing = [], []
    for da_n, da in data.items():
        for ins_id, current_ins in da.get("instances", {}).items():
            for edit_ins, input_pack in current_ins.items():
                item = {"da_n": da_n, "ins_id": ins_id, "edit_ins": edit_ins, **input_pack}
                path = expected_path(gen_dir, da_n, ins_id, edit_ins, make_dirs)
                if check_exist and osp.exists(path):
                    item["gen_img_path"] = path
                    existing.append(item)
                else:
                    cases.append(item)
    return cases, existing


def shard(cases: list, rank: int, world: int) -> list:
    return [cases[i] for i in ffdist.shard_indices(len(cases), rank, world)]


def merge_results(existing: list, per_rank: list) -> dict:
    final = list(existing)
    for res in per_rank:
        final.extend(res)
    out: dict = {}
    for item in final:
        out.setdefault(item["da_n"], {"instances": {}})["instances"].setdefault(item["ins_id"], {})[item["edit_ins"]] = item
    return out


def save_json(data: dict, path: str) -> None:
    with open(path, "w", encoding="utf-8") as f:
        json.dump(data, f, ensure_ascii=False, indent=4)


def read_and_resize_img(path: str, dsize=(512, 512)) -> np.ndarray:
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), dsize=dsize, interpolation=cv2.INTER_LANCZOS4)


def read_and_resize_mask(path: str, dsize=(512, 512)) -> np.ndarray:
    """uint8 0/1 [H,W] (the reference returns the same values replicated over 3 channels and reduces them later)."""
    import cv2
    m = cv2.imread(path)
    if m is None:
        raise FileNotFoundError(path)
    m = cv2.resize(m, dsize=dsize, interpolation=cv2.INTER_NEAREST)
    m[m > 0] = 1
    return np.ascontiguousarray(m[:, :, 0])


def save_img(img: np.ndarray, gen_dir: str, da_n, ins_id, edit_ins) -> str:
    import cv2
    path = expected_path(gen_dir, da_n, ins_id, edit_ins)
    cv2.imwrite(path, cv2.cvtColor(np.ascontiguousarray(img), cv2.COLOR_RGB2BGR))
    return path


def edit_param_2d(edit_param):
    """[dx,dy,dz,rx,ry,rz,sx,sy,sz] (annotation) -> (dx,dy,rz,sx,sy) (re_edit_2d, freefine_batch_infer_2d.py:29-31)."""
    dx, dy, _dz, _rx, _ry, rz, sx, sy, _sz = edit_param
    return float(dx), float(dy), float(rz), float(sx), float(sy)


def run(pipe, dst_base: str, edits_per_batch: int = 8, rank: int = 0, world: int = 1, generate=None, res: int = 512,
        settings: dict | None = None) -> dict | None:
    """Sharded GeoBench-2D inference.  `generate(images u8 [E,H,W,3], masks u8 [E,H,W], edit_params, inp_bgs, **settings)`
    defaults to pipe.FreeFine_generation_batch.  Every rank returns after the gather; rank 0 returns the merged dict
    (also written to <dst_base>/generated_results_freefine_2d.json), the others None."""
    import torch.distributed as tdist
    settings = dict(GEOBENCH_2D_SETTINGS if settings is None else settings)
    gen_dir = osp.join(dst_base, GEN_SUBDIR)
    if rank == 0:
        os.makedirs(gen_dir, exist_ok=True)
    with open(osp.join(dst_base, "annotations_2d.json"), "r", encoding="utf-8") as f:
        data = json.load(f)
    cases, existing = flatten_cases(data, gen_dir)
    mine = shard(cases, rank, world)
    if generate is None:
        generate = lambda imgs, masks, params, bgs, **kw: pipe.FreeFine_generation_batch(
            imgs, masks, params, [""] * len(params), inp_bgs=bgs, **kw)
    image_info = []
    for b0 in range(0, len(mine), edits_per_batch):
        batch = mine[b0:b0 + edits_per_batch]
        imgs = np.stack([read_and_resize_img(c["ori_img_path"], (res, res)) for c in batch])
        masks = np.stack([read_and_resize_mask(c["ori_mask_path"], (res, res)) for c in batch])
        bgs = np.stack([read_and_resize_img(osp.join(dst_base, INP_SUBDIR, str(c["da_n"]), str(c["ins_id"]), "inp_img.png"),
                                            (res, res)) for c in batch])
        out = generate(imgs, masks, [edit_param_2d(c["edit_param"]) for c in batch], bgs, **settings)
        for c, img in zip(batch, out):
            image_info.append({**c, "gen_img_path": save_img(np.asarray(img), gen_dir, c["da_n"], c["ins_id"], c["edit_ins"])})
    if world > 1:
        gathered = [None] * world
        tdist.all_gather_object(gathered, image_info)
    else:
        gathered = [image_info]
    if rank != 0:
        return None
    merged = merge_results(existing, gathered)
    save_json(merged, osp.join(dst_base, RESULT_JSON))
    return merged


# ---------------------------------------------------------------------------------------------------------------------
# GeoBench-3D (depth) driver: evaluation/FreeFine/freefine_batch_infer_3d_depth.py
# ---------------------------------------------------------------------------------------------------------------------
def run_3d_depth(pipe, dst_base: str, edits_per_batch: int = 4, rank: int = 0, world: int = 1, generate=None, res: int = 512,
                 settings: dict | None = None) -> dict | None:
    """main (:73-197) with E edits per stream batch: `annotations.json` flattens like the 2-D set (the driver's
    CustomDataset :27-66 is the same class), the coarse input is read from
    <dst_base>/coarse3d_depth_anything/<da_n>/<ins_id>/<edit_ins>.png (:120), the target mask from `target_mask_0`, the
    completion region from `draw_mask` (:122-123), the prompt is `obj_label` (:149); results go to
    Geo-Bench-3D/Gen_results_FreeFine_depth and generated_results_freefine_depth.json (:99, :191).
    `generate(images, masks, coarse_inputs, target_masks, draw_masks, prompts, **settings)` defaults to
    pipe.FreeFine_generation_batch (which skips its own warp when handed a coarse input)."""
    import torch.distributed as tdist
    settings = dict(GEOBENCH_3D_SETTINGS if settings is None else settings)
    gen_dir = osp.join(dst_base, GEN_SUBDIR_3D)
    if rank == 0:
        os.makedirs(gen_dir, exist_ok=True)
    with open(osp.join(dst_base, "annotations.json"), "r", encoding="utf-8") as f:
        data = json.load(f)
    cases, existing = flatten_cases(data, gen_dir)
    mine = shard(cases, rank, world)
    if generate is None:
        generate = lambda imgs, masks, coarse, tgt, draw, prompts, **kw: pipe.FreeFine_generation_batch(
            imgs, masks, None, prompts, coarse_inputs=coarse, target_masks=tgt, draw_masks=draw, **kw)
    image_info = []
    for b0 in range(0, len(mine), edits_per_batch):
        batch = mine[b0:b0 + edits_per_batch]
        rd_i = lambda pth: read_and_resize_img(pth, (res, res))
        rd_m = lambda pth: read_and_resize_mask(pth, (res, res))
        imgs = np.stack([rd_i(c["ori_img_path"]) for c in batch])
        masks = np.stack([rd_m(c["ori_mask_path"]) for c in batch])
        coarse = np.stack([rd_i(osp.join(dst_base, COARSE_SUBDIR_3D, str(c["da_n"]), str(c["ins_id"]), f"{c['edit_ins']}.png"))
                           for c in batch])
        tgt = np.stack([rd_m(c["target_mask_0"]) for c in batch])
        draw = np.stack([rd_m(c["draw_mask"]) for c in batch])
        out = generate(imgs, masks, coarse, tgt, draw, [c["obj_label"] for c in batch], **settings)
        for c, img in zip(batch, out):
            image_info.append({**c, "gen_img_path": save_img(np.asarray(img), gen_dir, c["da_n"], c["ins_id"], c["edit_ins"])})
    if world > 1:
        gathered = [None] * world
        tdist.all_gather_object(gathered, image_info)
    else:
        gathered = [image_info]
    if rank != 0:
        return None
    merged = merge_results(existing, gathered)
    save_json(merged, osp.join(dst_base, RESULT_JSON_3D))
    return merged


# ---------------------------------------------------------------------------------------------------------------------
# background generation for GeoBench-2D: evaluation/FreeFine/freefine_batch_infer_bggen_2d.py
# ---------------------------------------------------------------------------------------------------------------------
def inp_dir(dst_base: str, blending: bool) -> str:
    """:110-113 -- inp_img_blended is what the 2-D driver reads its backgrounds from (INP_SUBDIR)."""
    return osp.join(dst_base, "Geo-Bench-2D", "inp_img_blended" if blending else "inp_img_no_blend")


def expected_inp_path(inp_root: str, da_n, ins_id, make_dirs: bool = True) -> str:
    d = osp.join(inp_root, str(da_n), str(ins_id))
    if make_dirs:
        os.makedirs(d, exist_ok=True)
    return osp.join(d, "inp_img.png")


def flatten_cases_inpaint(data: dict, inp_root: str, check_exist: bool = True, make_dirs: bool = True):
    """CustomDatasetInpaint.__init__ (:41-66): ONE item per (image, instance) -- the input pack of its FIRST edit, without
    an `edit_ins` key -- resume on <inp_root>/<da_n>/<ins_id>/inp_img.png."""
    cases, existing = [], []
    for da_n, da in data.items():
        for ins_id, current_ins in da.get("instances", {}).items():
            first = next(iter(current_ins.keys())) if current_ins else None
            if first:
                item = {"da_n": da_n, "ins_id": ins_id, **current_ins[first]}
                path = expected_inp_path(inp_root, da_n, ins_id, make_dirs)
                if check_exist and osp.exists(path):
                    item["gen_img_path"] = path
                    existing.append(item)
                else:
                    cases.append(item)
    return cases, existing


def read_and_resize_mask_with_dilation(path: str, dsize=(512, 512), dilation_factor=None, forbit_area=None) -> np.ndarray:
    """src/utils/vis_utils.py:361-375: 3-channel {0,1} mask, nearest resize, optional k x k dilation (dilate_mask,
    :340-347 = cv2.dilate with a ones kernel), optional forbidden area zeroed."""
    import cv2
    m = cv2.imread(path)
    if m is None:
        raise FileNotFoundError(path)
    m = cv2.resize(m, dsize=tuple(dsize), interpolation=cv2.INTER_NEAREST)
    m[m > 0] = 1
    out = m
    if dilation_factor is not None:
        out = cv2.dilate(m.astype(np.uint8), np.ones((int(dilation_factor), int(dilation_factor)), np.uint8), iterations=1)
    if forbit_area is not None:
        out = np.where(forbit_area, 0, out)
    return out


def feather_blend(ori_img: np.ndarray, ori_mask: np.ndarray, generated: np.ndarray) -> np.ndarray:
    """The BrushNet-style paste of freefine_batch_infer_bggen_2d.py:184-188, literally: the (dilated, {0,1}, uint8,
    3-channel) mask is Gaussian-blurred with a 21 x 21 kernel IN uint8 and divided by 255, so the feather adds at most
    1/255 to the hard mask (the reference blurs a {0,1} mask where BrushNet blurs a {0,255} one); the generated image
    replaces the original inside it.  Result has the dtype of `generated`."""
    import cv2
    mask_blurred = cv2.GaussianBlur(ori_mask, (21, 21), 0) / 255
    mask_np = 1 - (1 - ori_mask) * (1 - mask_blurred)
    return (ori_img * (1 - mask_np) + generated * mask_np).astype(generated.dtype)


def save_inp_img(img: np.ndarray, inp_root: str, da_n, ins_id) -> str:
    import cv2
    path = expected_inp_path(inp_root, da_n, ins_id)
    cv2.imwrite(path, cv2.cvtColor(np.ascontiguousarray(img), cv2.COLOR_RGB2BGR))
    return path


def run_bggen_2d(pipe, dst_base: str, blending: bool = True, rank: int = 0, world: int = 1, generate=None, res: int = 512,
                 settings: dict | None = None, seed_fn=None) -> list:
    """main (:94-199): one background per (image, instance) with the object region (mask dilated by 30 px, :146) removed
    by FreeFine_background_generation under register_attention_control_4bggen, pasted back with `feather_blend` when
    `blending`, written to Geo-Bench-2D/inp_img_{blended,no_blend}/<da_n>/<ins_id>/inp_img.png.  The reference draws a
    fresh random seed per case (:162, "bring more diversity"); `seed_fn(case) -> int` makes runs repeatable (default: a
    hash of the case ids).  `generate(ori_img, ori_mask3, **settings) -> uint8 [H,W,3]` defaults to
    pipe.FreeFine_background_generation.  Returns this rank's list of {**case, inp_img_path} (the reference gathers an
    empty list, :196, and writes no JSON)."""
    import zlib
    settings = dict(BGGEN_2D_SETTINGS if settings is None else settings)
    root = inp_dir(dst_base, blending)
    if rank == 0:
        os.makedirs(root, exist_ok=True)
    with open(osp.join(dst_base, "annotations_2d.json"), "r", encoding="utf-8") as f:
        data = json.load(f)
    cases, _existing = flatten_cases_inpaint(data, root)
    if seed_fn is None:
        seed_fn = lambda c: zlib.crc32(f"{c['da_n']}/{c['ins_id']}".encode())
    if generate is None:
        # (reference quirk: the published driver also passes use_auto_draw / reduce_inp_artifacts, which
        # FreeFine_background_generation (model.py:1088) does not accept -- a TypeError as published; dropped here)
        drop = ("use_auto_draw", "reduce_inp_artifacts")
        generate = lambda img, mask3, **kw: pipe.FreeFine_background_generation(
            ori_img=img, ori_mask=mask3, **{k: v for k, v in kw.items() if k not in drop})
    done = []
    for c in shard(cases, rank, world):
        ori_img = read_and_resize_img(c["ori_img_path"], (res, res))
        ori_mask = read_and_resize_mask_with_dilation(c["ori_mask_path"], (res, res), dilation_factor=30, forbit_area=None)
        out = np.asarray(generate(ori_img, ori_mask, seed=int(seed_fn(c)), **settings))
        if blending:
            out = feather_blend(ori_img, ori_mask, out)
        done.append({**c, "inp_img_path": save_inp_img(out, root, c["da_n"], c["ins_id"])})
    return done
