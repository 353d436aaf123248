"""Host-side helpers of bench.py that need no GPU: the roofline `traffic` figure comes from the committed ncu capture and
only for the launch shape that capture was taken on; the workload description names the UNet-body mode."""
import argparse
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_is_read_from_the_committed_ncu_summary():
    t = bench.ncu_traffic(4096, 4096, 40, 32)
    lines = [l.split() for l in open(os.path.join(ROOT, "profiles", "r1_attn_ncu_summary.txt"))]
    rd = next(float(f[1]) for f in lines if f and f[0] == "dram__bytes_read.sum")
    wr = next(float(f[1]) for f in lines if f and f[0] == "dram__bytes_write.sum")
    assert abs(t["traffic"] - (rd + wr) * 1e6) < 1.0 and t["traffic_source"] == "profiles/r1_attn_ncu_summary.txt"
    assert bench.ncu_traffic(9216, 9216, 40, 16) == {}          # another launch shape: no claim


def test_workload_config_names_the_unet_body():
    a = argparse.Namespace(num_step=50, start_step=35, res=512, edits=8, preset="sd15", plain_unet=False)
    c = bench.workload_config(a)
    assert c["workload"].startswith("configs[1]") and "channels-last" in c["unet_body"]
    assert "15 inversion" in c["unet_calls_per_edit"]
    a.plain_unet = True
    assert "eager" in bench.workload_config(a)["unet_body"]
