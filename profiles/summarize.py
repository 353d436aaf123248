"""Turns the raw evidence of profiles/scripts/round1_profile.sh (gpurun_out/p_*) into the tracked summaries under
profiles/: launch-list shares and the `ncu --set full` metric summaries (read with `ncu -i ... --page raw --csv`)."""
import csv, io, os, re, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PRE = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out", os.environ.get("FF_PRE", "p_"))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"


def launches():
    rows = [r for r in csv.reader(l for l in open(PRE + "launches_bench.csv") if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}[r[ui]]
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        name = re.sub(r"(attn_masked_kv_kernel<[^>]*>).*", r"\1", name)
        name = re.sub(r"(attn_ring_kernel<[^>]*>).*", r"\1", name)
        name = re.sub(r"(attn_smallkv_kernel<[^>]*>).*", r"\1", name)
        name = name if "attn_masked" in name or "attn_ring" in name or "attn_smallkv" in name or "warp_affine" in name else re.sub(r"<.*", "", name)
        name = "at::layer_norm (eager)" if name.startswith("at::") and "layer_norm" in name else name
        tot[name][0] += v
        tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    n = sum(v[1] for v in tot.values())
    lines = ["# %s -- ncu launch list (gpu__time_duration.sum, --clock-control none) of:" % TAG,
             "#   python bench.py --steps 1 --warmup 0 --start-step 49 --no-cpu-baseline --no-e2e" + (" --no-extras" if TAG != "r1" else ""),
             "#   = one batch of 8 512x512 edits with 1 inversion UNet call (16 streams) + 1 TCA sampling call (32 streams)",
             "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  Raw list: %s_launches_bench.csv" % TAG,
             "# total %.1f ms over %d launches" % (total, n)]
    for name, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
        lines.append("%9.3f ms %5.1f%%  x%4d  %s" % (ms, 100 * ms / total, c, name[:90]))
    ours = {k: (round(v[0], 3), v[1]) for k, v in tot.items() if any(s in k for s in ("attn_masked", "attn_ring", "attn_smallkv", "gn_fused", "warp_affine", "ddim_", "mask_downsample", "cross_region", "kv_gather", "gn_stats", "gn_apply",
                                                                                      "geglu_kernel", "layer_norm_kernel", "layer_norm5_kernel", "bias_residual", "gn_small", "mask_prep",
                                                                                      "dilate_kernel", "upsample2x_nhwc", "concat_nhwc"))}
    lines.append("# our kernels (ms, launches): %s" % ours)
    lines.append("# share of our kernels: %.1f%%   share of all ff_attn_masked_kv launches: %.1f%%   ff_attn_plain_smallkv launches: %.1f%%" % (
        100 * sum(v[0] for v in ours.values()) / total, 100 * sum(v[0] for k, v in ours.items() if "attn_masked" in k or "attn_ring" in k) / total,
        100 * sum(v[0] for k, v in ours.items() if "attn_smallkv" in k) / total))
    open(os.path.join(OUT, TAG + "_launches_summary.txt"), "w").write("\n".join(lines) + "\n")
    os.replace(PRE + "launches_bench.csv", os.path.join(OUT, TAG + "_launches_bench.csv")) if False else None
    print("\n".join(lines[:14]))


WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def full(rep, out, header):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(r, units)))
    lines = list(header)
    for k in ["Kernel Name", "Grid Size", "Block Size"] + WANT:
        if k in d:
            lines.append("%-100s %s %s" % (k, d[k][0], d[k][1]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ix = {c: i for i, c in enumerate(h)}
    ops = defaultdict(int)
    tot = 0
    for row in rows[2:]:
        toks = [t for t in row[ix["Source"]].split() if not t.startswith("@")]
        if not toks:
            continue
        n = int(row[ix["Instructions Executed"]])
        ops[toks[0].split(".")[0]] += n
        tot += n
    lines.append("# SASS opcode mix (warp-level instructions executed, top 14 of %d):" % tot)
    for op, n in sorted(ops.items(), key=lambda kv: -kv[1])[:14]:
        lines.append("  %-12s %12d  %5.1f%%" % (op, n, 100.0 * n / tot))
    proof = sorted(o for o in ops if o.startswith(("UTC", "LDTM", "STTM", "UTMA", "LDGSTS")))
    lines.append("# tcgen05 / TMEM / TMA mnemonics present: %s" % ", ".join(proof))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    launches()
    if os.environ.get("FF_LAUNCHES_ONLY"):
        cp = os.path.join(OUT, TAG + "_launches_bench.csv")
        open(cp, "w").writelines(l for l in open(PRE + "launches_bench.csv") if not l.startswith("=="))
        sys.exit(0)
    timing = open(PRE + "attn_case_timing.txt").read().strip()
    if os.path.exists(PRE + "attn_d40_full.ncu-rep"):          # round 2: one capture per product attention kernel
        full(PRE + "attn_d40_full.ncu-rep", os.path.join(OUT, TAG + "_attn_ncu_summary.txt"),
             ["# %s -- ncu --set full --clock-control none, kernel attn_masked_kv_kernel<48,false> of profiles/attn_case.py" % TAG,
              "# (ff_attn_masked_kv, 8 batched edits = 32 streams x 8 heads, S=4096, d=40, 'tca' plans, synthetic masks, fp16 P.V path;",
              "#  the product kernel for d <= 40 in round 2: every ring-kernel layout measured slower at this shape)",
              "# CUDA-event timing of the same launches OUTSIDE ncu (profiles/attn_case.py 5):"] + ["#   " + l for l in timing.splitlines()])
        full(PRE + "attn_d80_full.ncu-rep", os.path.join(OUT, TAG + "_attn_d80_ncu_summary.txt"),
             ["# %s -- ncu --set full --clock-control none, kernel attn_ring_kernel<80,2,3,false> of profiles/attn_case.py" % TAG,
              "# (ff_attn_masked_kv, 32 streams x 8 heads, S=1024, d=80: the ring-buffered persistent kernel, csrc/attn_ring.cuh)",
              "# CUDA-event timing of the same launches OUTSIDE ncu (profiles/attn_case.py 5):"] + ["#   " + l for l in timing.splitlines()])
        sys.exit(0)
    full(PRE + "attn_full.ncu-rep", os.path.join(OUT, TAG + "_attn_ncu_summary.txt"),
         ["# %s -- ncu --set full --clock-control none, kernel attn_masked_kv_kernel<48,false> of profiles/attn_case.py" % TAG,
          "# (ff_attn_masked_kv, 8 batched edits = 32 streams x 8 heads, S=4096, d=40, 'tca' plans, synthetic masks, fp16 P.V path)",
          "# CUDA-event timing of the same launch OUTSIDE ncu (profiles/attn_case.py 5):"] + ["#   " + l for l in timing.splitlines()])
    full(PRE + "warp_full.ncu-rep", os.path.join(OUT, TAG + "_warp_ncu_summary.txt"),
         ["# %s -- ncu --set full --clock-control none, warp_affine_blend_kernel<float> of profiles/warp_case.py" % TAG,
          "# (N*C = 32768 channels of 64x64 fp32: 1.61 GB of algorithmic traffic, >> 126 MB L2); event timings: %s_hbm_kernels.json" % TAG])
