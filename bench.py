#!/usr/bin/env python
"""bench.py -- edits/s of the FreeFine hot path on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload (`config.workload`): BASELINE.json configs[1] -- SD1.5 2-D geometric edits (move + rotate + scale) at 512x512,
50-step DDIM schedule, 8 edits per stream batch per GPU, synthetic images / masks / transforms (freefine_b200/synth.py),
random-init SD1.5-shaped stand-in UNet (diffusers and the checkpoint are not available: no network).  A "step" is one
batch of 8 whole edits: coarse warp+blend -> DDIM inversion (UNet x n_inv, 16 streams) -> TCA sampling with local CFG
and masked DDPM steps (UNet x n_samp, 32 streams) -> decode.  Default n_inv = n_samp = 50 (`--start-step 0`); the
reference's GeoBench-2D default skips the first 35 schedule steps (`--start-step 35`, 15+15 UNet calls).

`value`   : whole-job edits/s with the inputs already resident in HBM (CUDA-event timed, max over ranks).
`e2e`     : the same through FreeFinePipeline.FreeFine_generation_batch with HOST buffers (pinned uint8 images and
            masks in, uint8 images out), copies inside the timed region.
`roofline`: the dominant kernel (ff_attn_masked_kv on the S=4096, d=40 TCA layers): algorithmic FLOPs per launch
            (freefine_b200.plans.algorithmic_flops) / mean launch duration from CUDA-event pairs recorded around every
            launch inside the timed region, against the measured bf16 peak of MEASURED_PEAKS.json.
`cpu_baseline`: the CPU restatement of the reference path (oracle/ff_pipeline_cpu.py, kind "port": the reference is
            Python that needs diffusers and cannot travel to the GPU box) timed on the host cores on a bounded sample
            (one inversion UNet step + one sampling UNet step of one 512x512 edit), extrapolated to a whole edit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "edits/s @512^2 50-step SD1.5 (FreeFine 2-D geometric edit)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edits", type=int, default=8, help="edits per stream batch per GPU")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--num-step", type=int, default=50)
    ap.add_argument("--start-step", type=int, default=0)
    ap.add_argument("--preset", default="sd15")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--channels-last", action="store_true", help="run the (library) UNet body in channels_last")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="A/B only: torch.backends.cudnn.benchmark = True")
    ap.add_argument("--plain-unet", action="store_true",
                    help="A/B only: eager NCHW UNet body instead of the channels-last fast path (csrc/unet_glue.cu)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1392.3), d.get("bf16_tflops", 1639.8), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated on the CPU (oracle port), bounded sample
# -------------------------------------------------------------------------------------------------------------------
def cpu_sample(args, n_samples=1):
    """Times one inversion UNet step (2 streams) + one TCA sampling step (4 streams) of one edit on the host cores.
    Returns (edits_per_s, seconds_per_sample list, cores)."""
    import numpy as np
    import torch
    from freefine_b200 import coarse_edit, synth
    from freefine_b200.standin import build_standin
    from oracle.ff_pipeline_cpu import OraclePipeline
    from oracle import ff_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    e = synth.make_edit(0, args.res)
    pipe = OraclePipeline(build_standin(args.preset))
    m = e["mask"]
    # coarse edit on the CPU: exact integer translation part of the edit (the warp is microseconds either way)
    dx, dy = int(round(e["edit_param"][0])), int(round(e["edit_param"][1]))
    tgt = np.roll(m, (dy, dx), (0, 1))
    coarse = np.where(tgt[:, :, None] != 0, np.roll(e["image"], (dy, dx), (0, 1)), e["image"])
    n_inv = n_samp = args.num_step - args.start_step
    times = []
    for _ in range(n_samples):
        t0 = time.perf_counter()
        inv = pipe.invert(coarse, e["image"], args.num_step, args.start_step, max_steps=1)
        t1 = time.perf_counter()
        inv_full = [inv[-1]] * (n_inv + 1)
        pipe.sample(inv_full, e["prompt"], tgt * 255, m, np.zeros_like(m), (args.res, args.res), args.num_step, args.start_step,
                    args.num_step, 7.5, 1.0, "tca", True, tgt * 255, True, 0.0, max_steps=1)
        t2 = time.perf_counter()
        times.append((t1 - t0, t2 - t1))
    t_inv = sum(t[0] for t in times) / len(times)
    t_samp = sum(t[1] for t in times) / len(times)
    return 1.0 / (n_inv * t_inv + n_samp * t_samp), times, cores, t_inv, t_samp


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, args.steps)
    warm = 1 if args.warmup > 0 else 0
    t0 = time.perf_counter()
    v, times, cores, t_inv, t_samp = cpu_sample(args, n_samples=warm + n)
    # drop the warm-up sample
    tt = times[warm:]
    t_inv = sum(t[0] for t in tt) / len(tt)
    t_samp = sum(t[1] for t in tt) / len(tt)
    n_calls = args.num_step - args.start_step
    v = 1.0 / (n_calls * (t_inv + t_samp))
    sample = (f"{warm} warm-up + {n} timed samples of [1 inversion UNet step (2 streams) + 1 TCA sampling step (4 streams)] "
              f"of one {args.res}x{args.res} edit, fp32, extrapolated x{n_calls} (t_inv={t_inv:.2f}s t_samp={t_samp:.2f}s)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "edits/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * args.edits / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "edits/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    _emit(line)


def ncu_traffic(sq, skv, d, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r1_attn_ncu_summary.txt: the same launch shape, S=4096 d=40 32 streams, run alone by profiles/attn_case.py);
    {} when the dominant launch of this run has another shape or the summary is absent."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_attn_ncu_summary.txt")
    if (sq, skv, d, B) != (4096, 4096, 40, 32) or not os.path.exists(path):
        return {}
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and f[2] in unit:
            tot += float(f[1]) * unit[f[2]]
    if tot <= 0:
        return {}
    return {"traffic": tot, "traffic_unit": "bytes per launch (DRAM read + write, ncu --set full)",
            "traffic_source": "profiles/r1_attn_ncu_summary.txt"}


def workload_config(args):
    n = args.num_step - args.start_step
    return {"workload": "configs[1]: SD1.5 2D geometric edit (move/rotate/scale) 512x512, 50-step, batch 8 per GPU",
            "resolution": args.res, "edits_per_step_per_gpu": args.edits, "num_step": args.num_step,
            "start_step": args.start_step, "unet_calls_per_edit": f"{n} inversion (2 streams) + {n} sampling (4 streams)",
            "method": "tca", "guidance_scale": 7.5, "eta": 1.0, "use_auto_draw": True, "reduce_inp_artifacts": True,
            "network": f"random-init SD1.5-shaped stand-in UNet ({args.preset}), bf16",
            "unet_body": "eager NCHW (A/B)" if getattr(args, "plain_unet", False) else
                         "channels-last fast path: cuDNN/cuBLAS + fused GroupNorm/SiLU, bias+residual, GEGLU, LayerNorm kernels",
            "l2": "working set per step (UNet weights 1.7 GB + activations) exceeds the 126 MB L2; no explicit flush",
            "parallelism": "independent edits, one model replica per GPU, no collective on the hot path"}


# -------------------------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            self.result = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from freefine_b200 import _lib, coarse_edit, ops, plans, synth
    from freefine_b200.pipeline import Attention_Modulator, FreeFinePipeline, register_attention_control
    from freefine_b200.standin import build_standin

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    if args.cudnn_benchmark:
        torch.backends.cudnn.benchmark = True

    if args.plain_unet:
        import freefine_b200.standin as _standin
        _standin.FAST_PATH = False
    parts = build_standin(args.preset, device=dev, dtype=torch.bfloat16)
    if args.channels_last:
        parts.unet.to(memory_format=torch.channels_last)
    controller = Attention_Modulator(start_layer=10)
    pipe = FreeFinePipeline.from_parts(parts, controller, device=dev)
    register_attention_control(pipe, controller)
    pipe.modify_unet_forward()

    E, R = args.edits, args.res
    kw = dict(guidance_scale=7.5, eta=1.0, end_step=args.num_step, num_step=args.num_step, start_step=args.start_step,
              method_type="tca", use_auto_draw=True, reduce_inp_artifacts=True, end_scale=0.0)
    total_steps = args.warmup + args.steps

    def batch(step):  # each rank edits its own images (weak scaling): edit index = ((step*world)+rank)*E + i
        return synth.make_batch((step * world + rank) * E, E, R)

    def thetas_for(b):
        return torch.tensor(np.stack([coarse_edit.cv2_theta(coarse_edit.edit_matrix(b["masks"][i], b["edit_params"][i]), R, R)
                                      for i in range(E)]), dtype=torch.float32)

    host = [batch(s) for s in range(total_steps)]
    for b in host:
        b["thetas"] = thetas_for(b)
        b["images_pin"] = torch.from_numpy(b["images"]).pin_memory()
        b["masks_pin"] = torch.from_numpy(b["masks"]).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_device(b_dev):
        out = pipe.FreeFine_generation_batch(b_dev["images"], b_dev["masks"], None, b_dev["prompts"], thetas=b_dev["thetas"], **kw)
        return out

    def run_host(b):
        return pipe.FreeFine_generation_batch(b["images_pin"], b["masks_pin"], b["edit_params"], b["prompts"], thetas=b["thetas"], **kw)

    def to_dev(b):
        return dict(images=b["images_pin"].to(dev), masks=b["masks_pin"].to(dev), thetas=b["thetas"].to(dev), prompts=b["prompts"])

    # ---- A: inputs resident in HBM -----------------------------------------------------------------------------
    dev_batches = [to_dev(b) for b in host]
    for s in range(args.warmup):
        run_device(dev_batches[s])
    barrier()
    ops.COUNTS.clear()
    prof = []
    ops.PROFILE = prof
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for s in range(args.warmup, total_steps):
            out = run_device(dev_batches[s])
        ev1.record()
        barrier()
    ops.PROFILE = None
    t_dev = max_over_ranks(ev0.elapsed_time(ev1) / 1e3)
    launches = int(sum(ops.COUNTS.values()))
    counts = dict(ops.COUNTS)
    value = world * E * args.steps / t_dev
    finite = bool(torch.isfinite(out.float()).all().item())

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------
    groups = {}
    for r in prof:
        ms = r["ev0"].elapsed_time(r["ev1"])
        key = (r["s_q"], r["s_kv"], r["d"], r["B"], id(r["plan"]))
        g = groups.setdefault(key, dict(ms=0.0, n=0, rec=r))
        g["ms"] += ms
        g["n"] += 1
    attn_ms_total = sum(g["ms"] for g in groups.values())
    # dominant shape group (by total time) -- all plans of that shape (the guidance weight changes per step)
    by_shape = {}
    for key, g in groups.items():
        sh = key[:4]
        b = by_shape.setdefault(sh, dict(ms=0.0, n=0, flops=0.0))
        rec = g["rec"]
        plan = rec["plan"]
        pop = rec["popcount"].cpu().numpy() if rec["popcount"] is not None else None
        fl = plans.algorithmic_flops(plan, rec["s_q"], rec["s_kv"], rec["d"], pop) if plan is not None else 0.0
        b["ms"] += g["ms"]
        b["n"] += g["n"]
        b["flops"] += fl * g["n"]
    dom = max(by_shape.items(), key=lambda kv: kv[1]["ms"])
    (sq, skv, d, B), dg = dom
    sustained, burst, peak_src = peaks()
    achieved = dg["flops"] / (dg["ms"] / 1e3) / 1e12 if dg["ms"] > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": f"attn_masked_kv_kernel (S_q={sq}, S_kv={skv}, d={d}, streams={B})",
                "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                "peak_source": peak_src + ", bf16_tflops_sustained (kernel timed inside a long step)",
                "frac_of_burst_peak": achieved / burst, "launches": dg["n"],
                "avg_launch_ms": dg["ms"] / max(dg["n"], 1),
                "algorithmic_gflop_per_launch": dg["flops"] / max(dg["n"], 1) / 1e9,
                "share_of_step": dg["ms"] / (t_dev * 1e3), "all_attention_share_of_step": attn_ms_total / (t_dev * 1e3),
                "traffic": None}
    roofline.update(ncu_traffic(sq, skv, d, B))

    # ---- B: end to end through the public API with host buffers ------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        run_host(host[0])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for s in range(args.warmup, total_steps):
            o = run_host(host[s])
        e1.record()
        barrier()
        t_host = max_over_ranks(max(e0.elapsed_time(e1) / 1e3, time.perf_counter() - t0))
        h2d = int(host[0]["images"].nbytes + host[0]["masks"].nbytes + host[0]["thetas"].numel() * 4)
        d2h = int(o.nbytes)
        e2e = {"value": world * E * args.steps / t_host, "unit": "edits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, times, cores, t_inv, t_samp = cpu_sample(args, n_samples=1)
        n = args.num_step - args.start_step
        cpu = {"value": v, "unit": "edits/s", "cores": cores, "kind": "port",
               "sample": f"1 inversion UNet step (2 streams, {t_inv:.1f}s) + 1 TCA sampling step (4 streams, {t_samp:.1f}s) of one "
                         f"{R}x{R} edit on the host cores (oracle/ff_pipeline_cpu.py, fp32), extrapolated x{n}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "edits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": workload_config(args), "clocks": clk.result,
                "e2e": e2e, "gpu_launches": launches, "gpu_launches_by_entry": counts, "roofline": roofline,
                "cpu_baseline": cpu, "output_finite": finite}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def _emit(line: dict):
    """The ONE JSON line of the contract, on the process's original stdout."""
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    global _JSON_OUT
    args = parse()
    # library chatter (e.g. the "NCCL version ..." banner some NCCL_DEBUG settings print on stdout) must not precede
    # the JSON line: file descriptor 1 is pointed at stderr for the whole run, the JSON goes to a dup of the real stdout
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
