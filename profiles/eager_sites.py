"""Which call sites still launch eager PyTorch kernels inside the timed edit?  One batch of 8 edits with the last two schedule
steps (2 inversion + 2 sampling UNet calls) under torch.profiler with Python stacks; prints, per aten op and innermost repo
source line, launches and device time.  (Diagnosis tool, not a bench: numbers taken under a profiler are never quoted.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import defaultdict
import torch
from torch.profiler import profile, ProfilerActivity
from freefine_b200 import selfcheck, synth

dev = torch.device("cuda:0")
pipe, _ = selfcheck.build_pipeline(dev, torch.bfloat16, preset="sd15")
b = synth.make_batch(0, 8, 512)
kw = dict(guidance_scale=7.5, eta=1.0, end_step=50, num_step=50, start_step=48, method_type="tca", end_scale=0.0)
for _ in range(2):
    pipe.FreeFine_generation_batch(b["images"], b["masks"], b["edit_params"], b["prompts"], **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    pipe.FreeFine_generation_batch(b["images"], b["masks"], b["edit_params"], b["prompts"], **kw)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for ev in prof.key_averages(group_by_stack_n=12):
    t = getattr(ev, "self_device_time_total", 0) or 0
    if t <= 0 or not ev.key.startswith("aten::"):
        continue
    site = next((s for s in (ev.stack or []) if "freefine_b200/" in s), "?")
    k = (ev.key, site.split("freefine_b200/")[-1][:80])
    agg[k][0] += ev.count
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
print(f"eager aten ops (self device time): {tot / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1] / 1e3:8.3f} ms x{v[0]:5d}  {k[0]:34s} {k[1]}")

print("\n# by input shapes (copy_ / add / cat / upsample / mul):")
agg2 = defaultdict(lambda: [0, 0.0])
for ev in prof.key_averages(group_by_input_shape=True):
    t = getattr(ev, "self_device_time_total", 0) or 0
    if t > 0 and ev.key in ("aten::copy_", "aten::add", "aten::cat", "aten::upsample_nearest2d", "aten::mul", "aten::add_", "aten::silu"):
        k = (ev.key, str(ev.input_shapes)[:110])
        agg2[k][0] += ev.count
        agg2[k][1] += t
for k, v in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v[1] / 1e3:8.3f} ms x{v[0]:5d}  {k[0]:26s} {k[1]}")
