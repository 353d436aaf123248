"""Builds freefine_b200/lib/libfreefine_b200.so (hand-written sm_100a CUDA behind the C ABI of include/freefine_b200.h).

    python -m freefine_b200.csrc.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The .so is built IN-TREE so that it travels to the GPU box with the
repo snapshot; it links the CUDA runtime statically and needs nothing from PyTorch.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libfreefine_b200.so")
SOURCES = ["ff_api.cu", "ddim_step.cu", "mask_pack.cu", "mask_prep.cu", "warp_blend.cu", "cross_blend.cu", "kv_prepare.cu", "unet_glue.cu", "linear_fused.cu", "attn_smallkv.cu", "attn_tcgen05.cu"]
HEADERS = ["ff_common.cuh", "attn_ring.cuh", "attn_t32.cuh", os.path.join("..", "..", "include", "freefine_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", extra_flags=()) -> str:
    """variant / extra_flags: experiment builds (e.g. variant="hilo", extra_flags=["-DFF_P_HILO"]) go to
    lib/libfreefine_b200_<variant>.so and are loaded with FREEFINE_B200_LIB=<path>; the product build has neither."""
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build", variant) if variant else os.path.join(HERE, "build")
    lib = os.path.join(LIB_DIR, f"libfreefine_b200_{variant}.so") if variant else LIB
    os.makedirs(obj_dir, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(HERE, h)) for h in HEADERS]
    objs, procs = [], []
    for src in SOURCES:
        sp = os.path.join(HERE, src)
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + hdrs):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                  (["-DFF_ENABLE_TRACE"] if os.environ.get("FF_TRACE") == "1" else []) + list(extra_flags) + \
                  ["-c", sp, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see messages above)")
    if force or procs or _stale(lib, objs):
        cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-lcublasLt"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return lib


if __name__ == "__main__":
    _variant = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), "")
    _extra = [a for a in sys.argv[1:] if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=_variant, extra_flags=_extra))
