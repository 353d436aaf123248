"""TEST INFRASTRUCTURE ONLY -- golden fixture for freefine_b200/geobench.py from the UNMODIFIED reference driver.

Imports evaluation/FreeFine/freefine_batch_infer_2d.py (read-only /root/reference, stubs of oracle/ref_import.py), runs
its CustomDataset (:88-131) on the synthetic annotation tree of `synthetic_annotations()` -- with two output PNGs already
present, to exercise the resume branch -- and its module-level re_edit_2d (:26-87) on one seeded image, and writes
tests/golden/geobench.json (paths relative to the scratch root).  Run in the build container:
    python -m oracle.make_golden_geobench
"""
import importlib.util
import json
import os
import os.path as osp
import sys
import tempfile

import numpy as np

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def synthetic_annotations(base: str) -> dict:
    """Three images, 1-3 instances each, 1-3 edits each, in a deliberately non-sorted key order; ids are strings as in
    the published annotations_2d.json."""
    data = {}
    k = 0
    for da_n, inst in (("12", {"3": 2, "0": 1}), ("7", {"1": 3}), ("105", {"2": 1, "5": 2, "4": 1})):
        data[da_n] = {"instances": {}, "src_img_path": f"{base}/src/{da_n}.png"}
        for ins_id, n_edit in inst.items():
            data[da_n]["instances"][ins_id] = {}
            for e in range(n_edit):
                k += 1
                data[da_n]["instances"][ins_id][str(e)] = {
                    "ori_img_path": f"{base}/src/{da_n}.png", "ori_mask_path": f"{base}/masks/{da_n}/{ins_id}.png",
                    "edit_param": [10.0 * k, -5.0 * k, 0, 0, 0, 3.0 * k, 1.0 + 0.01 * k, 1.0 + 0.01 * k, 1],
                    "edit_prompt": f"edit {k}", "obj_label": "thing"}
    return data


PRE_EXISTING = (("12", "3", "1"), ("105", "5", "0"))


def _load(name, *rel):
    spec = importlib.util.spec_from_file_location(name, osp.join(ref_import.REF_ROOT, *rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PRE_EXISTING_INP = (("7", "1"),)


def synthetic_metric_tree(base: str, rng_seed: int = 11):
    """PNG files + annotation dict for the WRAP_E golden: two images, one instance each, two / one samples."""
    import cv2
    rng = np.random.default_rng(rng_seed)
    data = {}
    for da_n, n in (("a", 2), ("b", 1)):
        data[da_n] = {"instances": {"0": {}}}
        for e in range(n):
            d = osp.join(base, da_n, str(e))
            os.makedirs(d, exist_ok=True)
            tgt = np.zeros((48, 40), np.uint8)
            tgt[10 + 3 * e:30, 8:25 + 2 * e] = 255
            paths = {}
            for nm, arr in (("coarse_input_path", rng.integers(0, 256, (48, 40, 3)).astype(np.uint8)),
                            ("gen_img_path", rng.integers(0, 256, (48, 40, 3)).astype(np.uint8)), ("tgt_mask_path", tgt)):
                paths[nm] = osp.join(d, nm + ".png")
                cv2.imwrite(paths[nm], arr)
            data[da_n]["instances"]["0"][str(e)] = paths
    return data


MD_PARAMS = ([7, -3, 0, 0, 0, 0, 1, 1, 1], [0, 0, 0, 0, 0, 25.0, 1, 1, 1], [0, 0, 0, 0, 0, 0, 1.3, 1.3, 1])


def md_mask():
    m = np.zeros((40, 56), np.float64)
    m[12:30, 20:44] = 1.0
    return m


def more_goldens() -> dict:
    """GeoBench-3D (depth) and background-generation drivers + metrics, from the unmodified reference modules."""
    import cv2
    out = {}
    d3 = _load("ref_batch_infer_3d_depth", "evaluation", "FreeFine", "freefine_batch_infer_3d_depth.py")
    bg = _load("ref_batch_infer_bggen_2d", "evaluation", "FreeFine", "freefine_batch_infer_bggen_2d.py")
    with tempfile.TemporaryDirectory() as base:
        rel = lambda items: [{k: (v.replace(base, "BASE") if isinstance(v, str) else v) for k, v in it.items()} for it in items]
        gen = osp.join(base, "Geo-Bench-3D", "Gen_results_FreeFine_depth")
        for da_n, ins_id, e in PRE_EXISTING:
            os.makedirs(osp.join(gen, da_n, ins_id), exist_ok=True)
            open(osp.join(gen, da_n, ins_id, f"{e}.png"), "wb").write(b"x")
        ds = d3.CustomDataset(synthetic_annotations("BASE"), gen)
        out["cases3d"], out["existing3d"] = rel(ds.cases), rel(ds.get_existing_results())
        inp = osp.join(base, "Geo-Bench-2D", "inp_img_blended")
        for da_n, ins_id in PRE_EXISTING_INP:
            os.makedirs(osp.join(inp, da_n, ins_id), exist_ok=True)
            open(osp.join(inp, da_n, ins_id, "inp_img.png"), "wb").write(b"x")
        di = bg.CustomDatasetInpaint(synthetic_annotations("BASE"), inp)
        out["inpaint_cases"], out["inpaint_existing"] = rel(di.cases), rel(di.get_existing_results())
        # read_and_resize_mask_with_dilation (vis_utils.py:361-375, imported by the bggen driver) on a 0/255 PNG
        m = np.zeros((100, 80), np.uint8)
        m[30:60, 20:50] = 255
        m[5, 70] = 255
        mp = osp.join(base, "m.png")
        cv2.imwrite(mp, m)
        dil = bg.read_and_resize_mask_with_dilation(mp, dsize=(64, 64), dilation_factor=30, forbit_area=None)
        plain = bg.read_and_resize_mask_with_dilation(mp, dsize=(64, 64))
        out["mask_dilation"] = {"shape": list(dil.shape), "sum": int(dil.sum()), "row20": dil[20, :, 0].tolist(),
                                "plain_sum": int(plain.sum()), "dtype": str(dil.dtype)}
        # WRAP_E (evaluation/metrics/wrap_error.py)
        we = _load("ref_wrap_error", "evaluation", "metrics", "wrap_error.py")
        out["wrap_e"] = float(we.calculate_we(synthetic_metric_tree(osp.join(base, "we")), "gen_img_path"))
    # MD: get_transform_coordinates (mean_distance.py:84-111); the package-relative import of the DIFT featurizer is stubbed
    import types
    pkg = types.ModuleType("MD")
    pkg.__path__ = [osp.join(ref_import.REF_ROOT, "evaluation", "metrics", "MD")]
    sys.modules["MD"] = pkg
    stub = types.ModuleType("MD.dift_sd")
    stub.SDFeaturizer = type("SDFeaturizer", (), {})
    sys.modules["MD.dift_sd"] = stub
    spec = importlib.util.spec_from_file_location("MD.mean_distance",
                                                  osp.join(ref_import.REF_ROOT, "evaluation", "metrics", "MD", "mean_distance.py"))
    md = importlib.util.module_from_spec(spec)
    sys.modules["MD.mean_distance"] = md
    spec.loader.exec_module(md)
    coords = []
    for prm in MD_PARAMS:
        c = md.get_transform_coordinates(list(prm), (40, 56), md_mask(), None)
        coords.append({"param": list(prm), "sum": float(c.sum()), "p_17_33": [float(c[17, 33, 0]), float(c[17, 33, 1])],
                       "p_0_55": [float(c[0, 55, 0]), float(c[0, 55, 1])]})
    out["md_coords"] = coords
    # inner loop of calculate_md (:156-166) on seeded feature maps, restated with the reference's own torch calls
    import torch
    g = torch.Generator().manual_seed(3)
    fs, fe = torch.randn(1, 24, 40, 56, generator=g), torch.randn(1, 24, 40, 56, generator=g)
    kps = [[13, 22], [20, 30], [29, 43], [15, 40]]
    fe[0, :, 19, 37] = fs[0, :, 20, 30] * 2.0            # an exact match somewhere else than the key point
    tc = md.get_transform_coordinates(list(MD_PARAMS[0]), (40, 56), md_mask(), None)
    cos = torch.nn.CosineSimilarity(dim=1)
    dists = []
    for k in kps:
        src_vec = fs[0, :, k[0], k[1]].view(1, 24, 1, 1)
        cos_map = cos(src_vec, fe).cpu().numpy()[0]
        max_rc = np.unravel_index(cos_map.argmax(), cos_map.shape)
        tp = torch.tensor(tc[k[0], k[1]])
        dists.append(float((tp - torch.tensor(max_rc)).float().norm()))
    out["md_dists"] = {"kps": kps, "dists": dists}
    return out


def main():
    ref_import.install_stubs()
    sys.path.insert(0, ref_import.REF_ROOT)
    spec = importlib.util.spec_from_file_location(
        "ref_batch_infer_2d", osp.join(ref_import.REF_ROOT, "evaluation", "FreeFine", "freefine_batch_infer_2d.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with tempfile.TemporaryDirectory() as base:
        gen = osp.join(base, "Geo-Bench-2D", "Gen_results_FreeFine_2d")
        for da_n, ins_id, e in PRE_EXISTING:
            os.makedirs(osp.join(gen, da_n, ins_id), exist_ok=True)
            open(osp.join(gen, da_n, ins_id, f"{e}.png"), "wb").write(b"x")
        data = synthetic_annotations("BASE")
        ds = mod.CustomDataset(data, gen)
        rel = lambda items: [{k: (v.replace(base, "BASE") if isinstance(v, str) else v) for k, v in it.items()} for it in items]
        out = {"cases": rel(ds.cases), "existing": rel(ds.get_existing_results())}
    # module-level re_edit_2d of the driver (the copy the benchmark actually runs)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, 64, 3)).astype(np.uint8)
    bg = rng.integers(0, 256, (64, 64, 3)).astype(np.uint8)
    mask = np.zeros((64, 64), np.uint8)
    mask[20:36, 24:44] = 1
    param = [6, -4, 0, 0, 0, 12.0, 1.1, 1.1, 1]
    coarse, tgt = mod.re_edit_2d(img, mask, list(param), bg)
    out["re_edit_2d"] = {"param": param, "coarse_sum": int(coarse.astype(np.int64).sum()), "tgt_sum": int(tgt.astype(np.int64).sum()),
                         "coarse_row40": coarse[40].tolist(), "tgt_row30": tgt[30].tolist()}
    out.update(more_goldens())
    dst = osp.join(ROOT, "tests", "golden", "geobench.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, len(out["cases"]), "cases,", len(out["existing"]), "existing")


if __name__ == "__main__":
    main()
