"""Achieved HBM bandwidth of the HBM-bound kernels (b) warp+blend and (c) CFG+DDIM step on L2-exceeding synthetic batches
(at real sizes -- 128 KiB of latents per edit -- these launches are latency-bound, so the roofline is
demonstrated at n_edits = 2048 / N*C = 32768).  CUDA-event timing of graph replays, 3 warm-ups, inputs >> 126 MB L2.
ALGORITHMIC bytes: (c) eps_u, eps_c, x, noise read + x_prev written (5 tensors of 2*4*h*w*4 B per edit) + 2*h*w mask bytes;
(c-inv) eps, x read + x_next written; (b) src read once + bg read + out written + mask bytes; K/V staging: K, V read,
K copy + padded fp16 V written + the int64 row index."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from freefine_b200 import ops

dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=10):
    """Per-call device time: `reps` calls captured in ONE CUDA graph and replayed (3 warm-up replays, best / median of 5
    timed replays, CUDA events on the replay stream), so that the Python / ctypes launch cost (~40 us per call, more than
    the run time of the smaller kernels) stays outside the measurement.  The buffers are far larger than L2 or, where
    stated, deliberately L2-sized."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    except Exception as e:          # not capturable: eager loop (launch cost included)
        sys.stderr.write(f"graph capture failed ({e}); eager timing\n")
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        return ms[0], ms[len(ms) // 2]
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / reps)
    ms.sort()
    del g
    return ms[0], ms[len(ms) // 2]


res = {}
E, h, w = 2048, 64, 64
eps4 = torch.randn(E, 4, 4, h, w, device=dev)
x = torch.randn(E, 2, 4, h, w, device=dev)
noise = torch.randn(E, 2, 4, h, w, device=dev)
cm = torch.randint(0, 3, (E, h, w), device=dev, dtype=torch.uint8)
vm = torch.randint(0, 3, (E, h, w), device=dev, dtype=torch.uint8)
out = torch.empty_like(x)
k = dict(sqrt_1m_at=0.6, sqrt_at=0.8, sqrt_ap=0.85, c_ddim=0.52, c_ddpm=0.5, sigma=0.14)
best, med = timeit(lambda: ops.ddim_cfg_step(eps4, x, noise, cm, vm, 7.5, out=out, **k))
byt = 5 * x.numel() * 4 + 2 * E * h * w
res["ddim_cfg_step"] = dict(n_edits=E, bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
eps = torch.randn(E, 2, 4, h, w, device=dev)
best, med = timeit(lambda: ops.ddim_inv_step(eps, x, 0.6, 0.8, 0.85, 0.52))
byt = 3 * x.numel() * 4
res["ddim_inv_step"] = dict(n=x.numel(), bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
del eps4, noise, eps
for NC, label in ((32768, "warp_blend_large"), (2048, "warp_blend_l2_resident")):
    src = torch.randn(1, NC, 64, 64, device=dev)
    bg = torch.randn(1, NC, 64, 64, device=dev)
    mask = (torch.rand(1, 64, 64, device=dev) > 0.5).to(torch.uint8)
    th = torch.tensor([[[0.95, 0.2, 0.05], [-0.2, 0.95, -0.03]]], device=dev)
    outb = torch.empty_like(bg)
    best, med = timeit(lambda: ops.warp_affine_blend(src, th, mask_src=mask, bg=bg, out=outb))
    byt = 3 * src.numel() * 4 + 2 * 64 * 64
    res[label] = dict(NC=NC, bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
    del src, bg, outb
# K/V staging: K gathered (bf16 copy), V gathered + converted into the padded fp16 layout (48 channels per 40)
heads, d, S, B = 8, 40, 4096, 64
kk = torch.randn(B, S, heads * d, device=dev).bfloat16()
vv = torch.randn(B, S, heads * d, device=dev).bfloat16()
idx = torch.randperm(B * S, device=dev)
best, med = timeit(lambda: ops.kv_gather_cast(kk, vv, heads, idx))
byt = 2 * kk.numel() * 2 + kk.numel() * 2 + B * S * heads * 48 * 2 + idx.numel() * 8
res["kv_gather_cast"] = dict(rows=B * S, bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak,
                             note="random row permutation (640-byte rows), output allocation inside the timed call")
del kk, vv
# UNet-body glue (csrc/unet_glue.cu) at the sizes of one 32-stream sampling call of the 512^2 batch (E = 8 edits):
# ALGORITHMIC bytes = each tensor once: GroupNorm x read + y written (the second read of x is meant to hit L2),
# bias+residual h, res read + out written, GEGLU [M,2F] read + [M,F] written, LayerNorm x read + y written.
def _nhwc(n, c, h, w):
    return torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)


for (n, c, hh, ww, G) in ((32, 320, 64, 64, 32), (32, 640, 32, 32, 32), (32, 960, 64, 64, 32)):
    xg = _nhwc(n, c, hh, ww)
    ga, be = torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16()
    add = torch.randn(n, c, device=dev)
    best, med = timeit(lambda: ops.group_norm_nhwc(xg, ga, be, G, 1e-5, add_nc=add, silu=True))
    byt = 2 * xg.numel() * 2
    res[f"group_norm_silu_nhwc_{n}x{c}x{hh}x{ww}"] = dict(bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6,
                                                          frac=byt / best / 1e6 / peak, note="2 launches (statistics, apply); output allocation inside the timed call")
    ref = lambda: F.silu(F.group_norm(xg, G, ga, be, 1e-5))
    b2, _ = timeit(ref)
    res[f"group_norm_silu_nhwc_{n}x{c}x{hh}x{ww}"]["eager_ms_best"] = b2
    if c == 320:
        r2 = _nhwc(n, c, hh, ww)
        best, med = timeit(lambda: ops.bias_residual_nhwc(xg, ga, r2))
        byt = 3 * xg.numel() * 2
        res["bias_residual_nhwc"] = dict(bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
        del r2
    del xg
hg = torch.randn(32, 4096, 2560, device=dev).bfloat16()
best, med = timeit(lambda: ops.geglu(hg))
byt = hg.numel() * 2 * 3 // 2
res["geglu_32x4096x2560"] = dict(bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
xa, ga2 = hg.chunk(2, dim=-1)
b2, _ = timeit(lambda: xa * F.gelu(ga2))
res["geglu_32x4096x2560"]["eager_ms_best"] = b2
del hg, xa, ga2
xl = torch.randn(32, 4096, 320, device=dev).bfloat16()
ga, be = torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16()
best, med = timeit(lambda: ops.layer_norm(xl, ga, be, 1e-5))
byt = 2 * xl.numel() * 2
res["layer_norm_32x4096x320"] = dict(bytes=byt, ms_best=best, ms_median=med, gbs=byt / best / 1e6, frac=byt / best / 1e6 / peak)
b2, _ = timeit(lambda: F.layer_norm(xl, (320,), ga, be, 1e-5))
res["layer_norm_32x4096x320"]["eager_ms_best"] = b2
del xl
res["peak_hbm_gbs"] = peak
print(json.dumps(res, indent=1))
