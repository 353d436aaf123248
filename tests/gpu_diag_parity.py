"""GPU diagnostic (not a test): whole-edit parity numbers of every golden edit on both UNet paths -- fp32 (plain PyTorch
UNet body, TF32 off) and bf16 (channels-last fast path, the path bench.py times).  Prints one JSON object per line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import selfcheck

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = np.load(os.path.join(selfcheck.GOLDEN_DIR, "pipeline.npz"))
names = sorted({k.split("/")[0] for k in g.files if k.endswith("/params_json")})
only = sys.argv[1:]
for dt_name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    for name in names:
        if only and name not in only:
            continue
        pipe, _ = selfcheck.build_pipeline(dev, dt)
        t0 = time.time()
        r = selfcheck.run_golden_edit(pipe, name, g)
        r.pop("edit_img")
        r.update(unet=dt_name, seconds=round(time.time() - t0, 2))
        print(json.dumps(r), flush=True)
    if not only or "config1" in only:
        pipe, _ = selfcheck.build_pipeline(dev, dt)
        t0 = time.time()
        r = selfcheck.run_config1(pipe)
        r.update(unet=dt_name, seconds=round(time.time() - t0, 2))
        print(json.dumps(r), flush=True)
