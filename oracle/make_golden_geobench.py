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
    dst = osp.join(ROOT, "tests", "golden", "geobench.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, len(out["cases"]), "cases,", len(out["existing"]), "existing")


if __name__ == "__main__":
    main()
