"""Roofline measurements of the individual kernels, callable from bench.py after its timed region and
from profiles/hbm_kernels.py, so that every fraction quoted in DESIGN.md is reproduced by the driver's own bench run.

* HBM-bound kernels -- (b) warp+blend, (c) CFG+DDIM step / inversion step, K/V staging, the channels-last UNet-body glue:
  L2-exceeding synthetic batches (at real sizes, 128 KiB of latents per edit, these launches are latency-bound), per-call
  device time from CUDA-graph replays (3 warm-up replays, best of 5 timed replays of `reps` calls each, CUDA events on
  the replay stream) so that the Python / ctypes launch cost stays outside; achieved = ALGORITHMIC bytes / time.
  ALGORITHMIC bytes: every tensor the operation must read or write, once (DESIGN.md section 4 per kernel).
* attention -- one launch per SD1.5 layer shape with the TCA plan of 8 batched edits (4 at 768^2) and synthetic
  GeoBench-like masks; achieved = ALGORITHMIC FLOPs (plans.algorithmic_flops) / time.
* gpu_active_frac -- sum of kernel durations / wall span of one step, from the CUPTI activity records torch.profiler
  collects: what fraction of a step the GPU is busy (the rest is host launch gaps).
"""
from __future__ import annotations

import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import ops, plans, synth


def _timeit(fn, reps=10):
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
        return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
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
    del g
    return min(ms)


def _row(kernel, byt, ms, peak, **extra):
    gbs = byt / ms / 1e6
    return dict(kernel=kernel, bound="hbm", bytes=int(byt), ms=ms, gbs=gbs, achieved=gbs, peak=peak, unit="GB/s", frac=gbs / peak,
                traffic=None, **extra)


def _nhwc(n, c, h, w, dev):
    return torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)


def hbm_rooflines(dev, peak_gbs: float, with_eager: bool = False):
    """-> list of roofline rows for the HBM-bound kernels.  ~3 GB of scratch at most at a time."""
    rows = []
    E, h, w = 2048, 64, 64
    eps4 = torch.randn(E, 4, 4, h, w, device=dev)
    x = torch.randn(E, 2, 4, h, w, device=dev)
    noise = torch.randn(E, 2, 4, h, w, device=dev)
    cm = torch.randint(0, 3, (E, h, w), device=dev, dtype=torch.uint8)
    vm = torch.randint(0, 3, (E, h, w), device=dev, dtype=torch.uint8)
    out = torch.empty_like(x)
    k = dict(sqrt_1m_at=0.6, sqrt_at=0.8, sqrt_ap=0.85, c_ddim=0.52, c_ddpm=0.5, sigma=0.14)
    ms = _timeit(lambda: ops.ddim_cfg_step(eps4, x, noise, cm, vm, 7.5, out=out, **k))
    rows.append(_row("ff_ddim_cfg_step", 5 * x.numel() * 4 + 2 * E * h * w, ms, peak_gbs, size=f"n_edits={E}, 64x64 latents"))
    eps = torch.randn(E, 2, 4, h, w, device=dev)
    ms = _timeit(lambda: ops.ddim_inv_step(eps, x, 0.6, 0.8, 0.85, 0.52))
    rows.append(_row("ff_ddim_inv_step", 3 * x.numel() * 4, ms, peak_gbs, size=f"{x.numel()} elements"))
    del eps4, noise, eps, x, out, cm, vm
    NC = 32768
    src = torch.randn(1, NC, 64, 64, device=dev)
    bg = torch.randn(1, NC, 64, 64, device=dev)
    mask = (torch.rand(1, 64, 64, device=dev) > 0.5).to(torch.uint8)
    th = torch.tensor([[[0.95, 0.2, 0.05], [-0.2, 0.95, -0.03]]], device=dev)
    outb = torch.empty_like(bg)
    ms = _timeit(lambda: ops.warp_affine_blend(src, th, mask_src=mask, bg=bg, out=outb))
    rows.append(_row("ff_warp_affine_blend", 3 * src.numel() * 4 + 2 * 64 * 64, ms, peak_gbs, size=f"N*C={NC}, 64x64 fp32"))
    del src, bg, outb
    heads, d, S, B = 8, 40, 4096, 64
    kk = torch.randn(B, S, heads * d, device=dev).bfloat16()
    vv = torch.randn(B, S, heads * d, device=dev).bfloat16()
    idx = torch.randperm(B * S, device=dev)
    ms = _timeit(lambda: ops.kv_gather_cast(kk, vv, heads, idx))
    byt = 2 * kk.numel() * 2 + kk.numel() * 2 + B * S * heads * 48 * 2 + idx.numel() * 8
    rows.append(_row("ff_kv_gather_cast", byt, ms, peak_gbs, size=f"{B} streams x {S} tokens, random row permutation"))
    del kk, vv, idx
    # (the last two are the small levels of a call: register-resident single-read kernel; L2-resident tensors)
    for (n, c, hh, ww, G) in ((32, 320, 64, 64, 32), (32, 960, 64, 64, 32), (32, 1280, 16, 16, 32), (32, 2560, 8, 8, 32)):
        xg = _nhwc(n, c, hh, ww, dev)
        ga, be = torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16()
        add = torch.randn(n, c, device=dev)
        ms = _timeit(lambda: ops.group_norm_nhwc(xg, ga, be, G, 1e-5, add_nc=add, silu=True))
        r = _row("ff_group_norm_nhwc", 2 * xg.numel() * 2, ms, peak_gbs, size=f"{n}x{c}x{hh}x{ww} bf16, +[N,C] addend, SiLU")
        if with_eager:
            r["eager_ms"] = _timeit(lambda: F.silu(F.group_norm(xg, G, ga, be, 1e-5)))
        rows.append(r)
        if c == 320:
            r2 = _nhwc(n, c, hh, ww, dev)
            ms = _timeit(lambda: ops.bias_residual_nhwc(xg, ga, r2))
            rows.append(_row("ff_bias_residual_nhwc", 3 * xg.numel() * 2, ms, peak_gbs, size=f"{n}x{c}x{hh}x{ww} bf16"))
            del r2
        del xg
    xu = _nhwc(32, 640, 32, 32, dev)
    ms = _timeit(lambda: ops.upsample2x_nhwc(xu))
    r = _row("ff_upsample2x_nhwc", 5 * xu.numel() * 2, ms, peak_gbs, size="32x640x32x32 -> 64x64 bf16")
    if with_eager:
        r["eager_ms"] = _timeit(lambda: F.interpolate(xu, scale_factor=2.0, mode="nearest"))
    rows.append(r)
    ca_, cb_ = _nhwc(32, 640, 64, 64, dev), _nhwc(32, 320, 64, 64, dev)
    ms = _timeit(lambda: ops.concat_nhwc(ca_, cb_))
    r = _row("ff_concat_nhwc", 2 * (ca_.numel() + cb_.numel()) * 2, ms, peak_gbs, size="32x(640+320)x64x64 bf16")
    if with_eager:
        r["eager_ms"] = _timeit(lambda: torch.cat([ca_, cb_], dim=1))
    rows.append(r)
    del xu, ca_, cb_
    hg = torch.randn(32, 4096, 2560, device=dev).bfloat16()
    ms = _timeit(lambda: ops.geglu(hg))
    r = _row("ff_geglu", hg.numel() * 2 * 3 // 2, ms, peak_gbs, size="32x4096x2560 bf16")
    if with_eager:
        xa, ga2 = hg.chunk(2, dim=-1)
        r["eager_ms"] = _timeit(lambda: xa * F.gelu(ga2))
    rows.append(r)
    del hg
    # text cross-attention of a 32-stream call (77 keys): HBM-bound on Q in + O out (K/V stay in shared memory)
    for (bq, sq, dq) in ((32, 4096, 40), (32, 1024, 80), (32, 256, 160)):
        qq = torch.randn(bq, sq, 8 * dq, device=dev).bfloat16()
        kq_ = torch.randn(bq, 77, 8 * dq, device=dev).bfloat16()
        vq_ = torch.randn(bq, 77, 8 * dq, device=dev).bfloat16()
        ms = _timeit(lambda: ops.attn_plain_smallkv(qq, kq_, vq_, 8, dq ** -0.5))
        rows.append(_row("ff_attn_plain_smallkv", 2 * qq.numel() * 2 + 2 * kq_.numel() * 2, ms, peak_gbs,
                         size=f"{bq} streams x {sq} queries x 77 keys, 8 heads, d={dq}, bf16"))
        del qq, kq_, vq_
    xl = torch.randn(32, 4096, 320, device=dev).bfloat16()
    ga, be = torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16()
    ms = _timeit(lambda: ops.layer_norm(xl, ga, be, 1e-5))
    r = _row("ff_layer_norm", 2 * xl.numel() * 2, ms, peak_gbs, size="32x4096x320 bf16")
    if with_eager:
        r["eager_ms"] = _timeit(lambda: F.layer_norm(xl, (320,), ga, be, 1e-5))
    rows.append(r)
    del xl
    torch.cuda.empty_cache()
    return rows


def attention_case(dev, S: int, d: int, E: int = 8, heads: int = 8, res: int = 512, method: str = "tca", cg: float = 0.5):
    """Inputs of one TCA layer call of E batched edits (4E streams) at S tokens: q, sorted K, staged V, plan, bit-vectors."""
    hw = int(round(S ** 0.5))
    g = torch.Generator().manual_seed(0)
    q = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    k = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    v = torch.randn(4 * E, S, heads * d, generator=g).to(dev).bfloat16()
    masks = []
    for e in range(E):
        b = synth.make_edit(e, res)
        sh = (res * 20 // 512, -(res * 30 // 512))
        masks += [b["mask"], np.roll(b["mask"], sh, (0, 1))]                      # src (fg_ref), tgt (fg_retain)
    bits, pop = ops.mask_downsample_pack(torch.from_numpy(np.stack(masks)).to(dev), hw, hw)
    plan_np = plans.tca_plan(E, heads, method, cg, lambda e: 2 * e, lambda e: 2 * e + 1, prefix=True)
    plan = ops.to_device_bytes(plan_np, dev)
    shifts = torch.arange(32, device=dev, dtype=torch.int32)
    key_bits = ((bits[:, :, None] >> shifts) & 1).reshape(bits.shape[0], -1)[:, :S]
    idx = plans.kv_sort_index(key_bits, [2 * (s // 4) if s % 2 else -1 for s in range(4 * E)])
    k, v = ops.kv_gather_cast(k, v, heads, idx)
    flops = plans.algorithmic_flops(plan_np, S, S, d, pop.cpu().numpy())
    return dict(q=q, k=k, v=v, plan=plan, bits=bits, pop=pop, heads=heads, scale=d ** -0.5, flops=flops)


def attention_rooflines(dev, peak_tflops: float, shapes=((4096, 40, 8, 512), (1024, 80, 8, 512), (256, 160, 8, 512), (9216, 40, 4, 768),
                                                         (2304, 80, 4, 768)), reps=5):
    """One ff_attn_masked_kv launch per layer shape (S, d, edits, resolution), timed alone with CUDA events (best of
    `reps`, 2 warm-ups; the operands of the S >= 1024 shapes exceed L2 only partly: K/V are meant to be L2 hits) ->
    roofline rows against the burst bf16 peak (a kernel timed in isolation)."""
    rows = []
    for S, d, E, res in shapes:
        c = attention_case(dev, S, d, E, res=res)
        out = ops.attn_masked_kv(c["q"], c["k"], c["v"], c["plan"], c["heads"], c["scale"], c["bits"], c["pop"])
        ops.attn_masked_kv(c["q"], c["k"], c["v"], c["plan"], c["heads"], c["scale"], c["bits"], c["pop"], out=out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            ops.attn_masked_kv(c["q"], c["k"], c["v"], c["plan"], c["heads"], c["scale"], c["bits"], c["pop"], out=out)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        tf = c["flops"] / (ms * 1e-3) / 1e12
        rows.append(dict(kernel="ff_attn_masked_kv", bound="tensor", size=f"S={S} d={d} streams={4 * E} (TCA, {E} edits at {res}^2)",
                         flops=c["flops"], ms=ms, achieved=tf, peak=peak_tflops, unit="TFLOP/s", frac=tf / peak_tflops, traffic=None))
        del c, out
    torch.cuda.empty_cache()
    return rows


def gpu_active_frac(fn):
    """Runs fn() once under torch.profiler (CUPTI kernel activity records) -> (sum of kernel durations / span from the
    first kernel start to the last kernel end, number of kernels, span in ms); None when the profiler is unavailable."""
    try:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        ks = [e for e in prof.events() if getattr(e, "device_type", None) is not None and str(e.device_type).endswith("CUDA")
              and e.time_range is not None]
        if not ks:
            return None
        t0 = min(e.time_range.start for e in ks)
        t1 = max(e.time_range.end for e in ks)
        busy = sum(e.time_range.end - e.time_range.start for e in ks)
        return dict(frac=float(busy) / float(t1 - t0), kernels=len(ks), span_ms=float(t1 - t0) / 1e3,
                    note="sum of kernel durations / (last kernel end - first kernel start) of one step, CUPTI via torch.profiler; "
                         "kernels on one stream do not overlap")
    except Exception as e:      # pragma: no cover
        return dict(frac=None, error=str(e)[:200])
