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
