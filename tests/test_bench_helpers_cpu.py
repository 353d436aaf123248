"""Host-side helpers of bench.py that need no GPU: the roofline `traffic` figure comes from the committed ncu capture and
only for the launch shape that capture was taken on; the workload description names the UNet-body mode."""
import argparse
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_traffic_is_read_from_the_committed_ncu_summary():
    t = bench.ncu_traffic(4096, 4096, 40, 32)
    name = next(f for f in ("r2b_attn_ncu_summary.txt", "r2_attn_ncu_summary.txt", "r1_attn_ncu_summary.txt") if os.path.exists(os.path.join(ROOT, "profiles", f)))
    lines = [l.split() for l in open(os.path.join(ROOT, "profiles", name))]
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = next(float(f[1]) * unit[f[2]] for f in lines if f and f[0] == "dram__bytes_read.sum")
    wr = next(float(f[1]) * unit[f[2]] for f in lines if f and f[0] == "dram__bytes_write.sum")
    assert abs(t["traffic"] - (rd + wr)) < 1.0 and t["traffic_source"].startswith("profiles/" + name)
    assert bench.ncu_traffic(9216, 9216, 40, 16) == {}          # another launch shape: no claim


def test_workload_config_names_the_unet_body():
    a = argparse.Namespace(num_step=50, start_step=35, res=512, edits=8, preset="sd15", plain_unet=False)
    c = bench.workload_config(a)
    assert c["workload"].startswith("configs[1]") and "channels-last" in c["unet_body"]
    assert "15 inversion" in c["unet_calls_per_edit"]
    a.plain_unet = True
    assert "eager" in bench.workload_config(a)["unet_body"]
    a.res, a.unet_dtype = 768, "fp32"
    c = bench.workload_config(a)
    assert "768x768" in c["workload"] and c["workload"].startswith("configs[4]") and "fp32" in c["network"]
    assert "fp32" in bench.workload_config(a, reference=True)["network"]
