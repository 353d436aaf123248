"""Pins the CPU whole-edit restatement (oracle/ff_pipeline_cpu.py, the `port` used as bench.py's CPU baseline)
against latents produced by the UNMODIFIED reference pipeline (tests/golden/pipeline.npz)."""
import numpy as np
import pytest
import torch

from freefine_b200.standin import build_standin
from oracle import cases
from oracle.ff_pipeline_cpu import OraclePipeline


@pytest.mark.parametrize("name", list(cases.PIPE_CASES))
def test_oracle_pipeline_matches_reference(golden, name):
    g = golden["pipeline"]
    c = cases.PIPE_CASES[name]
    pipe = OraclePipeline(build_standin("tiny"), noise_fn=lambda k, shape: cases.step_noise(c["seed"], k, shape))
    inv = pipe.invert(g[name + "/coarse"], g[name + "/img"], c["num_step"], c["start_step"])
    inv_ref = torch.from_numpy(g[name + "/inverted"])
    assert len(inv) == len(inv_ref)
    assert float((inv[-1] - inv_ref[-1]).norm() / inv_ref[-1].norm()) < 1e-5
    ori_mask = g[name + "/ori_mask"][:, :, 0]
    lat = pipe.sample(inv, c["prompt"], g[name + "/tgt_mask"], ori_mask, g[name + "/draw"], (c["res"], c["res"]), c["num_step"],
                      c["start_step"], c["end_step"], c["gs"], c["eta"], c["method"], c["use_auto_draw"], g[name + "/cons"],
                      c["reduce_inp_artifacts"], c["end_scale"])
    ref = torch.from_numpy(g[name + "/latents"])
    assert len(lat) == len(ref)
    rel = float((lat[-1] - ref[-1]).norm() / ref[-1].norm())
    assert rel < 1e-4, rel
    img = pipe.decode(lat[-1])
    assert np.abs(img[0].astype(int) - g[name + "/edit_img"].astype(int)).max() <= 1
