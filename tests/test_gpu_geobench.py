"""GeoBench drivers end to end on the GPU (-m gpu): synthetic annotation trees on disk, the real pipeline (tiny stand-in
network, 128x128) behind geobench.run / run_3d_depth / run_bggen_2d -- the callers of the hot path.
Checked: the reference's output trees and JSON files appear, every image is what the per-edit entry point of the reference
surface produces for the same inputs, resume skips finished cases, WRAP_E runs on the result."""
import json
import os
import os.path as osp

import cv2
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RES = 128
STEPS = dict(num_step=6, end_step=6)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from freefine_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _tree(base, three_d):
    """Two images x (2, 1) instances x (2, 1 | 1) edits, PNG inputs written like the benchmark's."""
    from freefine_b200 import synth
    data, k = {}, 0
    for da_n, inst in (("3", {"0": 2, "1": 1}), ("11", {"2": 1})):
        e0 = synth.make_edit(40 + len(data), RES)
        os.makedirs(osp.join(base, "src"), exist_ok=True)
        cv2.imwrite(osp.join(base, "src", f"{da_n}.png"), cv2.cvtColor(e0["image"], cv2.COLOR_RGB2BGR))
        data[da_n] = {"instances": {}}
        for ins_id, n_edit in inst.items():
            m = synth.make_edit(60 + k, RES)["mask"]
            os.makedirs(osp.join(base, "masks", da_n), exist_ok=True)
            cv2.imwrite(osp.join(base, "masks", da_n, f"{ins_id}.png"), m * 255)
            d = osp.join(base, "Geo-Bench-2D", "inp_img_blended", da_n, ins_id)
            os.makedirs(d, exist_ok=True)
            cv2.imwrite(osp.join(d, "inp_img.png"), cv2.cvtColor(synth.make_edit(80 + k, RES)["image"], cv2.COLOR_RGB2BGR))
            data[da_n]["instances"][ins_id] = {}
            for e in range(n_edit):
                k += 1
                pack = {"ori_img_path": osp.join(base, "src", f"{da_n}.png"),
                        "ori_mask_path": osp.join(base, "masks", da_n, f"{ins_id}.png"),
                        "edit_param": [6.0 * k, -4.0 * k, 0, 0, 0, 5.0 * k, 1.0 + 0.02 * k, 1.0 + 0.02 * k, 1],
                        "edit_prompt": f"edit {k}", "obj_label": "thing"}
                if three_d:
                    cd = osp.join(base, "coarse3d_depth_anything", da_n, ins_id)
                    os.makedirs(cd, exist_ok=True)
                    cv2.imwrite(osp.join(cd, f"{e}.png"), cv2.cvtColor(synth.make_edit(100 + k, RES)["image"], cv2.COLOR_RGB2BGR))
                    pack["target_mask_0"] = osp.join(cd, f"{e}_tgt.png")
                    pack["draw_mask"] = osp.join(cd, f"{e}_draw.png")
                    cv2.imwrite(pack["target_mask_0"], np.roll(m, (7, -5), (0, 1)) * 255)
                    cv2.imwrite(pack["draw_mask"], np.roll(m, (7, 9), (0, 1)) * 255)
                data[da_n]["instances"][ins_id][str(e)] = pack
    json.dump(data, open(osp.join(base, "annotations.json" if three_d else "annotations_2d.json"), "w"))
    return data


def _rgb(p):
    return cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB)


def test_run_2d_matches_per_edit_entry_point(dev, tmp_path):
    from freefine_b200 import coarse_edit, geobench, metrics, selfcheck
    base = str(tmp_path)
    data = _tree(base, three_d=False)
    pipe, _ = selfcheck.build_pipeline(dev, torch.float32)
    settings = dict(geobench.GEOBENCH_2D_SETTINGS, start_step=2, **STEPS)
    merged = geobench.run(pipe, base, edits_per_batch=3, res=RES, settings=settings)
    assert json.load(open(osp.join(base, geobench.RESULT_JSON))) == merged
    n, we_data = 0, {}
    for da_n, da in data.items():
        for ins_id, edits in da["instances"].items():
            for e, pack in edits.items():
                item = merged[da_n]["instances"][ins_id][e]
                assert item["gen_img_path"] == osp.join(base, geobench.GEN_SUBDIR, da_n, ins_id, f"{e}.png")
                got = _rgb(item["gen_img_path"])
                # the reference's own sequence for one case: re_edit_2d on the host, then FreeFine_generation
                img = geobench.read_and_resize_img(pack["ori_img_path"], (RES, RES))
                mask = geobench.read_and_resize_mask(pack["ori_mask_path"], (RES, RES))
                bg = geobench.read_and_resize_img(osp.join(base, geobench.INP_SUBDIR, da_n, ins_id, "inp_img.png"), (RES, RES))
                coarse, tgt = coarse_edit.re_edit_2d(img, mask, geobench.edit_param_2d(pack["edit_param"]), bg)[:2]
                want = pipe.FreeFine_generation(img, mask, coarse, tgt, "", draw_mask=None, cons_area=tgt,
                                                **{k: v for k, v in settings.items()})
                assert got.shape == (RES, RES, 3)
                assert np.abs(got.astype(int) - want.astype(int)).mean() < 1.0, (da_n, ins_id, e)
                # WRAP_E inputs: coarse input, generated image, target mask
                d = osp.join(base, "we", da_n, ins_id)
                os.makedirs(d, exist_ok=True)
                cv2.imwrite(osp.join(d, f"{e}_c.png"), cv2.cvtColor(coarse, cv2.COLOR_RGB2BGR))
                cv2.imwrite(osp.join(d, f"{e}_t.png"), (np.asarray(tgt) > 0).astype(np.uint8) * 255)
                we_data.setdefault(da_n, {"instances": {}})["instances"].setdefault(ins_id, {})[e] = {
                    "coarse_input_path": osp.join(d, f"{e}_c.png"), "gen": item["gen_img_path"], "tgt_mask_path": osp.join(d, f"{e}_t.png")}
                n += 1
    assert n == 4
    we = metrics.calculate_we(we_data, "gen")
    assert np.isfinite(we) and 0.0 <= we <= 1.0
    # resume: nothing left to do, the JSON is rebuilt from the files on disk
    again = geobench.run(pipe, base, edits_per_batch=3, res=RES, settings=settings,
                         generate=lambda *a, **k: (_ for _ in ()).throw(AssertionError("nothing should be generated")))
    assert again == merged


def test_run_3d_depth_matches_per_edit_entry_point(dev, tmp_path):
    from freefine_b200 import geobench, selfcheck
    base = str(tmp_path)
    data = _tree(base, three_d=True)
    pipe, _ = selfcheck.build_pipeline(dev, torch.float32)
    settings = dict(geobench.GEOBENCH_3D_SETTINGS, start_step=2, **STEPS)
    merged = geobench.run_3d_depth(pipe, base, edits_per_batch=3, res=RES, settings=settings)
    assert json.load(open(osp.join(base, geobench.RESULT_JSON_3D))) == merged
    rd_i = lambda p: geobench.read_and_resize_img(p, (RES, RES))
    rd_m = lambda p: geobench.read_and_resize_mask(p, (RES, RES))
    for da_n, da in data.items():
        for ins_id, edits in da["instances"].items():
            for e, pack in edits.items():
                item = merged[da_n]["instances"][ins_id][e]
                got = _rgb(item["gen_img_path"])
                coarse = rd_i(osp.join(base, geobench.COARSE_SUBDIR_3D, da_n, ins_id, f"{e}.png"))
                tgt, draw = rd_m(pack["target_mask_0"]), rd_m(pack["draw_mask"])
                want = pipe.FreeFine_generation(rd_i(pack["ori_img_path"]), rd_m(pack["ori_mask_path"]), coarse, tgt,
                                                pack["obj_label"], draw_mask=draw, cons_area=tgt, **settings)
                assert np.abs(got.astype(int) - want.astype(int)).mean() < 1.0, (da_n, ins_id, e)


def test_run_bggen_2d_writes_backgrounds(dev, tmp_path):
    from freefine_b200 import geobench
    from freefine_b200.pipeline import Attention_Modulator, FreeFinePipeline, register_attention_control_4bggen
    from freefine_b200.standin import build_standin
    base = str(tmp_path)
    data = _tree(base, three_d=False)
    for root, _dirs, files in os.walk(osp.join(base, "Geo-Bench-2D")):          # start without any background
        for f in files:
            os.remove(osp.join(root, f))
    controller = Attention_Modulator(start_layer=10)
    pipe = FreeFinePipeline.from_parts(build_standin("tiny", device=dev), controller, device=dev)
    register_attention_control_4bggen(pipe, controller)                         # the driver's registration (:103)
    pipe.modify_unet_forward()
    settings = dict(geobench.BGGEN_2D_SETTINGS, num_step=6, end_step=4, start_step=1)
    done = geobench.run_bggen_2d(pipe, base, blending=True, res=RES, settings=settings)
    assert len(done) == sum(len(da["instances"]) for da in data.values()) == 3
    for it in done:
        img = _rgb(it["inp_img_path"])
        assert img.shape == (RES, RES, 3) and it["inp_img_path"].startswith(osp.join(base, geobench.INP_SUBDIR))
        ori = geobench.read_and_resize_img(it["ori_img_path"], (RES, RES))
        m3 = geobench.read_and_resize_mask_with_dilation(it["ori_mask_path"], (RES, RES), dilation_factor=30)
        outside = m3[:, :, 0] == 0
        # the feather paste keeps the original outside the (dilated) object region, up to its 1/255 leak + uint8 truncation
        assert np.abs(img[outside].astype(int) - ori[outside].astype(int)).max() <= 2
        assert np.isfinite(img.astype(np.float64)).all()
    assert geobench.run_bggen_2d(pipe, base, blending=True, res=RES, settings=settings) == []     # resume
