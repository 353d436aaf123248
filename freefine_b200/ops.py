"""Torch-tensor front end of the C ABI (include/freefine_b200.h).  PyTorch is only the owner of device memory and
streams here: every function validates its tensors, takes raw pointers and calls libfreefine_b200.so on
torch.cuda.current_stream().  No function has a PyTorch / CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import FF_DT_BF16, FF_DT_F32

_DT = {torch.float32: FF_DT_F32, torch.bfloat16: FF_DT_BF16}

# Measurement hooks (bench.py): COUNTS = kernel launches issued through the C ABI, by entry point;
# PROFILE = None or a list that receives one CUDA-event pair per ff_attn_masked_kv launch (events are recorded on
# the launching stream, so the pair brackets exactly that kernel).
COUNTS: dict = {}
PROFILE = None
PLAN_REGISTRY: dict = {}     # device plan pointer -> host (numpy) plan, filled by the controller; read only when PROFILE is on


# NVTX ranges around the phases of an edit (inversion / sampling steps, UNet calls, attention launches, step kernels) so
# that an nsys timeline of the hot path is one command: FREEFINE_NVTX=1 nsys profile python bench.py ...  (SURVEY.md 5).
# Off by default: a push/pop pair per attention launch is ~1 us of host time on a path that issues 3000 launches per step.
NVTX = os.environ.get("FREEFINE_NVTX", "0") == "1"


@contextlib.contextmanager
def nvtx_range(name: str):
    if NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


LIBRARY_ENTRIES = frozenset({"ff_linear_bias_residual"})     # C-ABI entries that launch library code (cuBLASLt), not our kernels


def _count(name: str):
    COUNTS[name] = COUNTS.get(name, 0) + 1


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> C.c_void_p:
    # (raw handle of torch's current stream on the current device: the C call is ~4x cheaper than building a
    # torch.cuda.Stream object, and this runs once per kernel launch)
    if _RAW_STREAM is not None:
        return C.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype, name: str, dims=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (freefine_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dims is not None and t.dim() != dims:
        raise ValueError(f"{name} must have {dims} dims, got {tuple(t.shape)}")
    return t


# ---------------------------------------------------------------------------------------------------------------
# (c) DDIM / CFG steps
# ---------------------------------------------------------------------------------------------------------------
def ddim_cfg_step(eps4, x, noise, cfg_mask, var_mask, guidance_scale, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm,
                  sigma, want_pred_x0=False, out=None):
    """eps4 [E,4,C,h,w] (or [4E,C,h,w]), x [E,2,C,h,w] (or [2E,C,h,w]) fp32; masks u8 [E,h,w]; returns x_prev like x
    (and pred_x0).  See ff_ddim_cfg_step in include/freefine_b200.h; reference model.py:605-611,134-198."""
    _chk(eps4, torch.float32, "eps4")
    _chk(x, torch.float32, "x")
    Cc, h, w = x.shape[-3:]
    n_edits = x.numel() // (2 * Cc * h * w)
    if eps4.numel() != 2 * x.numel():
        raise ValueError(f"eps4 {tuple(eps4.shape)} must hold 4 streams per edit for x {tuple(x.shape)}")
    _chk(var_mask, torch.uint8, "var_mask")
    if var_mask.numel() != n_edits * h * w:
        raise ValueError("var_mask must be [n_edits,h,w]")
    if cfg_mask is not None:
        _chk(cfg_mask, torch.uint8, "cfg_mask")
        if cfg_mask.numel() != n_edits * h * w:
            raise ValueError("cfg_mask must be [n_edits,h,w]")
    if noise is not None:
        _chk(noise, torch.float32, "noise")
        if noise.numel() != x.numel():
            raise ValueError("noise must have the shape of x")
    x_prev = torch.empty_like(x) if out is None else _chk(out, torch.float32, "out")
    x0 = torch.empty_like(x) if want_pred_x0 else None
    lib = _lib.load()
    rc = lib.ff_ddim_cfg_step(_ptr(eps4), _ptr(x), _ptr(noise), _ptr(cfg_mask), _ptr(var_mask), guidance_scale,
                              sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma, _ptr(x_prev), _ptr(x0), n_edits,
                              Cc, h, w, _stream())
    _lib.check(rc, "ff_ddim_cfg_step")
    _count("ff_ddim_cfg_step")
    return (x_prev, x0) if want_pred_x0 else x_prev


def ddim_step(eps2, x, noise, var_mask, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma, want_pred_x0=False):
    """ctrl_step alone (model.py:134-198) on an already guidance-combined eps2 [E,2,C,h,w]."""
    _chk(eps2, torch.float32, "eps2")
    _chk(x, torch.float32, "x")
    if eps2.shape != x.shape:
        raise ValueError("eps2 and x must have the same shape")
    Cc, h, w = x.shape[-3:]
    n_edits = x.numel() // (2 * Cc * h * w)
    _chk(var_mask, torch.uint8, "var_mask")
    if var_mask.numel() != n_edits * h * w:
        raise ValueError("var_mask must be [n_edits,h,w]")
    if noise is not None:
        _chk(noise, torch.float32, "noise")
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_pred_x0 else None
    rc = _lib.load().ff_ddim_step(_ptr(eps2), _ptr(x), _ptr(noise), _ptr(var_mask), sqrt_1m_at, sqrt_at, sqrt_ap,
                                  c_ddim, c_ddpm, sigma, _ptr(x_prev), _ptr(x0), n_edits, Cc, h, w, _stream())
    _lib.check(rc, "ff_ddim_step")
    _count("ff_ddim_step")
    return (x_prev, x0) if want_pred_x0 else x_prev


def ddim_cfg_step_compose(eps, streams_per_edit, x, noise, cfg_mask, var_mask, guidance_scale, sqrt_1m_at, sqrt_at,
                          sqrt_ap, c_ddim, c_ddpm, sigma):
    """Composition step (model.py:418-431): eps [E*(N+2),C,h,w] = [e, r_1..r_N, c_e] per edit, x [E,C,h,w] (edit stream
    only); guidance from streams 0 and N+1, single-stream masked DDPM/DDIM update."""
    _chk(eps, torch.float32, "eps")
    _chk(x, torch.float32, "x")
    Cc, h, w = x.shape[-3:]
    E = x.numel() // (Cc * h * w)
    if eps.numel() != E * streams_per_edit * Cc * h * w:
        raise ValueError("eps must hold streams_per_edit streams per edit")
    _chk(var_mask, torch.uint8, "var_mask")
    if cfg_mask is not None:
        _chk(cfg_mask, torch.uint8, "cfg_mask")
    if noise is not None:
        _chk(noise, torch.float32, "noise")
    x_prev = torch.empty_like(x)
    rc = _lib.load().ff_ddim_cfg_step_compose(_ptr(eps), streams_per_edit, _ptr(x), _ptr(noise), _ptr(cfg_mask), _ptr(var_mask),
                                              guidance_scale, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma,
                                              _ptr(x_prev), None, E, Cc, h, w, _stream())
    _lib.check(rc, "ff_ddim_cfg_step_compose")
    _count("ff_ddim_cfg_step_compose")
    return x_prev


def ddim_inv_step(eps, x, sqrt_1m_at, sqrt_at, sqrt_an, c_next, want_pred_x0=False):
    """reference inv_step, model.py:109-132."""
    _chk(eps, torch.float32, "eps")
    _chk(x, torch.float32, "x")
    if eps.shape != x.shape:
        raise ValueError("eps and x must have the same shape")
    x_next = torch.empty_like(x)
    x0 = torch.empty_like(x) if want_pred_x0 else None
    rc = _lib.load().ff_ddim_inv_step(_ptr(eps), _ptr(x), sqrt_1m_at, sqrt_at, sqrt_an, c_next, _ptr(x_next), _ptr(x0),
                                      x.numel(), _stream())
    _lib.check(rc, "ff_ddim_inv_step")
    _count("ff_ddim_inv_step")
    return (x_next, x0) if want_pred_x0 else x_next


# ---------------------------------------------------------------------------------------------------------------
# (b) warp + blend
# ---------------------------------------------------------------------------------------------------------------
def warp_affine_blend(src, theta, dsize=None, mask_src=None, bg=None, mode="bilinear", want_mask=False, out=None):
    """src [N,C,H,W] f32/bf16, theta [N,2,3] (or [2,3]) f32 normalised (param2theta), dsize (width, height).
    With mask_src [N,H,W] u8 + bg [N,C,dH,dW]: out = warped_mask ? warped_src : bg (re_edit_2d blend)."""
    if src.dtype not in _DT:
        raise TypeError(f"src must be float32 or bfloat16, got {src.dtype}")
    _chk(src, src.dtype, "src", 4)
    N, Cc, H, W = src.shape
    dW, dH = (W, H) if dsize is None else dsize
    theta = theta.to(device=src.device, dtype=torch.float32)
    if theta.dim() == 2:
        theta = theta[None].expand(N, 2, 3)
    theta = theta.contiguous()
    if theta.shape != (N, 2, 3):
        raise ValueError(f"theta must be [N,2,3], got {tuple(theta.shape)}")
    if mask_src is not None:
        _chk(mask_src, torch.uint8, "mask_src")
        if mask_src.numel() != N * H * W:
            raise ValueError("mask_src must be [N,H,W]")
        if bg is None:
            raise ValueError("mask_src needs a background")
        _chk(bg, src.dtype, "bg", 4)
        if bg.shape != (N, Cc, dH, dW):
            raise ValueError("bg must be [N,C,dH,dW]")
    if out is None:
        out = torch.empty((N, Cc, dH, dW), dtype=src.dtype, device=src.device)
    else:
        _chk(out, src.dtype, "out", 4)
        if out.shape != (N, Cc, dH, dW):
            raise ValueError("out must be [N,C,dH,dW]")
    mask_out = torch.empty((N, dH, dW), dtype=torch.uint8, device=src.device) if (want_mask and mask_src is not None) else None
    rc = _lib.load().ff_warp_affine_blend(_ptr(src), _ptr(theta), _ptr(mask_src), _ptr(bg), _ptr(out), _ptr(mask_out),
                                          N, Cc, H, W, dH, dW, 0 if mode == "bilinear" else 1, _DT[src.dtype],
                                          _stream())
    _lib.check(rc, "ff_warp_affine_blend")
    _count("ff_warp_affine_blend")
    return (out, mask_out) if want_mask else out


# ---------------------------------------------------------------------------------------------------------------
# masks
# ---------------------------------------------------------------------------------------------------------------
def mask_words(S: int) -> int:
    return (S + 31) // 32


def mask_downsample_pack(masks, h, w, bits=None, popcount=None):
    """masks u8 [n,H,W] -> (bits u32-as-int32 [n, words], popcount int32 [n]) at the h x w token grid
    (process_mask_before_attention, attention.py:841-855)."""
    _chk(masks, torch.uint8, "masks", 3)
    n, H, W = masks.shape
    words = mask_words(h * w)
    if bits is None:
        bits = torch.empty((n, words), dtype=torch.int32, device=masks.device)
    if popcount is None:
        popcount = torch.empty((n,), dtype=torch.int32, device=masks.device)
    rc = _lib.load().ff_mask_downsample_pack(_ptr(masks), n, H, W, h, w, _ptr(bits), bits.shape[1], _ptr(popcount),
                                             _stream())
    _lib.check(rc, "ff_mask_downsample_pack")
    _count("ff_mask_downsample_pack")
    return bits, popcount


def mask_prep(shifted, ori, draw, cons, lat_hw, use_auto_draw, reduce_inp_artifacts):
    """prepare_various_mask (model.py:1432-1512) for E edits in one launch: uint8 [E,H,W] masks (any non-zero = set; draw /
    cons may be None where the reference does not read them) -> (fg, shifted, ori) uint8 [E,H,W] in {0,1} and
    (completion, local_var) uint8 [E,h,w] in {0,1,2} (quirk Q1).  See ff_mask_prep."""
    _chk(shifted, torch.uint8, "shifted", 3)
    _chk(ori, torch.uint8, "ori", 3)
    E, H, W = shifted.shape
    for name, t in (("ori", ori), ("draw", draw), ("cons", cons)):
        if t is not None:
            _chk(t, torch.uint8, name, 3)
            if tuple(t.shape) != (E, H, W):
                raise ValueError(f"{name} must be [{E},{H},{W}], got {tuple(t.shape)}")
    h, w = lat_hw
    dev = shifted.device
    fg, sh, orib = (torch.empty((E, H, W), dtype=torch.uint8, device=dev) for _ in range(3))
    comp, lvar = (torch.empty((E, h, w), dtype=torch.uint8, device=dev) for _ in range(2))
    rc = _lib.load().ff_mask_prep(_ptr(shifted), _ptr(ori), _ptr(draw), _ptr(cons), E, H, W, h, w, int(bool(use_auto_draw)),
                                  int(bool(reduce_inp_artifacts)), _ptr(fg), _ptr(sh), _ptr(orib), _ptr(comp), _ptr(lvar), _stream())
    _lib.check(rc, "ff_mask_prep")
    _count("ff_mask_prep")
    return fg, sh, orib, comp, lvar


def dilate_mask(mask, k):
    """cv2.dilate(mask != 0, ones(k, k)) on uint8 [N,H,W] (or [H,W]) CUDA masks -> {0,1}.  See ff_dilate_mask."""
    m = mask if mask.dim() == 3 else mask[None]
    _chk(m, torch.uint8, "mask", 3)
    out = torch.empty_like(m)
    rc = _lib.load().ff_dilate_mask(_ptr(m), _ptr(out), m.shape[0], m.shape[1], m.shape[2], int(k), _stream())
    _lib.check(rc, "ff_dilate_mask")
    _count("ff_dilate_mask")
    return out if mask.dim() == 3 else out[0]


# ---------------------------------------------------------------------------------------------------------------
# (a) attention
# ---------------------------------------------------------------------------------------------------------------
class StagedV:
    """Staging of V for ff_attn_masked_kv (ff_kv_gather_cast): data [Bk, Skv, heads, v_head_stride] fp16 or bf16 with a
    ones column at channel head_dim."""
    __slots__ = ("data", "heads", "head_dim")

    def __init__(self, data, heads, head_dim):
        self.data, self.heads, self.head_dim = data, heads, head_dim

    def values(self):
        """[Bk, Skv, heads*head_dim] copy of the real channels (tests)."""
        return self.data[..., : self.head_dim].reshape(self.data.shape[0], self.data.shape[1], -1)


P_OPERAND_DTYPE = {"f16": torch.float16, "bf16x2": torch.bfloat16}


def kv_gather_cast(k, v, heads, row_index=None, gather_k=True, p_operand="f16"):
    """K/V staging in one pass: rows gathered by `row_index` (int64 over the flattened [Bk*Skv] rows, None = identity),
    V written into the padded per-head layout the kernel reads -- as fp16 (p_operand "f16": one fp16 P operand, the
    fast path; saturating) or bf16 ("bf16x2": hi+lo bf16 P pair).  Returns (k_sorted bf16, StagedV); K is returned
    untouched when there is no gather."""
    _chk(k, torch.bfloat16, "k", 3)
    _chk(v, torch.bfloat16, "v", 3)
    if v.shape != k.shape or k.shape[2] % heads:
        raise ValueError("k and v must have the same shape [Bk,Skv,heads*d]")
    Bk, Sk, Ck = k.shape
    d = Ck // heads
    if row_index is not None:
        _chk(row_index, torch.int64, "row_index", 1)
        if row_index.numel() != Bk * Sk:
            raise ValueError("row_index must hold one source row per K/V row")
    lib = _lib.load()
    vhs = lib.ff_attn_v_head_stride(d)
    do_k = gather_k and row_index is not None
    k_out = torch.empty_like(k) if do_k else None
    vdt = P_OPERAND_DTYPE[p_operand]
    v_out = torch.empty((Bk, Sk, heads, vhs), dtype=vdt, device=v.device)
    rc = lib.ff_kv_gather_cast(_ptr(k) if do_k else None, _ptr(v), _ptr(row_index), _ptr(k_out), _ptr(v_out),
                               _lib.FF_DT_F16 if vdt == torch.float16 else FF_DT_BF16, Bk * Sk, heads, d, _stream())
    _lib.check(rc, "ff_kv_gather_cast")
    _count("ff_kv_gather_cast")
    return (k_out if do_k else k), StagedV(v_out, heads, d)


def attn_masked_kv(q, k, v, plan, heads, scale, bitmasks=None, popcount=None, out_dtype=None, out=None,
                   p_operand="f16"):
    """q [B,Sq,C], k [Bk,Skv,C] bf16; v a StagedV (kv_gather_cast) or a bf16 [Bk,Skv,C] tensor, which is staged here
    according to `p_operand`; plan: uint8 CUDA tensor holding FFAttnHeadPlan[B*heads] (plans.to_device);
    bitmasks int32 [n_masks, words], popcount int32 [n_masks].  Returns [B,Sq,C] in out_dtype (default bf16)."""
    _chk(q, torch.bfloat16, "q", 3)
    _chk(k, torch.bfloat16, "k", 3)
    B, Sq, Cc = q.shape
    Bk, Skv, Ck = k.shape
    if Ck != Cc or Cc % heads:
        raise ValueError(f"bad shapes q{tuple(q.shape)} k{tuple(k.shape)} heads={heads}")
    if not isinstance(v, StagedV):
        _chk(v, torch.bfloat16, "v", 3)
        if v.shape != k.shape:
            raise ValueError(f"bad shapes k{tuple(k.shape)} v{tuple(v.shape)}")
        v = kv_gather_cast(k, v, heads, None, p_operand=p_operand)[1]
    vt = v.data
    _chk(vt, vt.dtype if vt.dtype in (torch.float16, torch.bfloat16) else torch.float16, "v", 4)
    if tuple(vt.shape[:3]) != (Bk, Skv, heads) or v.head_dim * heads != Ck:
        raise ValueError(f"staged V {tuple(vt.shape)} does not match k{tuple(k.shape)} heads={heads}")
    _chk(plan, torch.uint8, "plan")
    if plan.numel() != B * heads * _lib.PLAN_BYTES:
        raise ValueError(f"plan holds {plan.numel()} bytes, expected {B * heads * _lib.PLAN_BYTES}")
    out_dtype = torch.bfloat16 if out_dtype is None else out_dtype
    if out is None:
        out = torch.empty((B, Sq, Cc), dtype=out_dtype, device=q.device)
    else:
        _chk(out, out_dtype, "out", 3)
    a = _lib.FFAttnArgs()
    a.q, a.k, a.v, a.out, a.plan = q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), plan.data_ptr()
    if bitmasks is not None:
        _chk(bitmasks, torch.int32, "bitmasks", 2)
        _chk(popcount, torch.int32, "popcount", 1)
        a.bitmasks, a.mask_popcount = bitmasks.data_ptr(), popcount.data_ptr()
        a.n_masks, a.mask_words = bitmasks.shape[0], bitmasks.shape[1]
    else:
        a.bitmasks, a.mask_popcount, a.n_masks, a.mask_words = None, None, 0, 0
    a.n_streams, a.n_kv_streams, a.heads, a.head_dim = B, Bk, heads, Cc // heads
    a.s_q, a.s_kv = Sq, Skv
    a.out_dtype = _DT[out_dtype]
    a.scale = float(scale)
    a.v_dtype = _lib.FF_DT_F16 if vt.dtype == torch.float16 else FF_DT_BF16
    a.v_head_stride = vt.shape[3]
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    with nvtx_range(f"ff_attn_masked_kv S_q={Sq} S_kv={Skv} d={Cc // heads} streams={B}"):
        rc = _lib.load().ff_attn_masked_kv(C.byref(a), _stream())
    _lib.check(rc, "ff_attn_masked_kv")
    COUNTS["ff_attn_masked_kv"] = COUNTS.get("ff_attn_masked_kv", 0) + 1
    if prof is not None:
        ev1.record()
        prof.append(dict(ev0=ev0, ev1=ev1, B=B, Bk=Bk, s_q=Sq, s_kv=Skv, d=Cc // heads, heads=heads, plan=PLAN_REGISTRY.get(plan.data_ptr()),
                         popcount=popcount))
    return out


SMALLKV_MAX_KEYS = 256      # (128 for head_dim 8)
SMALLKV_HEAD_DIMS = (8, 40, 80, 160)


def attn_plain_smallkv(q, k, v, heads, scale, out_dtype=None):
    """softmax(q k^T * scale) v, stream i over K/V stream i, for s_kv <= 256 and head_dim in SMALLKV_HEAD_DIMS: q [B,Sq,C],
    k / v [B,Skv,C] bf16 (v unstaged).  Returns [B,Sq,C] in out_dtype (default bf16).  See ff_attn_plain_smallkv."""
    _chk(q, torch.bfloat16, "q", 3)
    _chk(k, torch.bfloat16, "k", 3)
    _chk(v, torch.bfloat16, "v", 3)
    B, Sq, Cc = q.shape
    if k.shape[0] != B or k.shape[2] != Cc or v.shape != k.shape or Cc % heads:
        raise ValueError(f"bad shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} heads={heads}")
    out_dtype = torch.bfloat16 if out_dtype is None else out_dtype
    out = torch.empty((B, Sq, Cc), dtype=out_dtype, device=q.device)
    with nvtx_range(f"ff_attn_plain_smallkv S_q={Sq} S_kv={k.shape[1]} d={Cc // heads} streams={B}"):
        rc = _lib.load().ff_attn_plain_smallkv(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, heads, Cc // heads, Sq, k.shape[1],
                                               float(scale), _DT[out_dtype], _stream())
    _lib.check(rc, "ff_attn_plain_smallkv")
    _count("ff_attn_plain_smallkv")
    return out


def cross_region_blend(hs, bitmasks, region_ids):
    """hs [4E,S,C] (f32/bf16) in place: c_e <- region ? c_e : u_e, c_r <- u_r  (attention.py:1381-1383)."""
    if hs.dtype not in _DT:
        raise TypeError("hs must be float32 or bfloat16")
    _chk(hs, hs.dtype, "hs", 3)
    B, S, Cc = hs.shape
    if B % 4:
        raise ValueError("hs must hold 4 streams per edit")
    _chk(bitmasks, torch.int32, "bitmasks", 2)
    _chk(region_ids, torch.int32, "region_ids", 1)
    rc = _lib.load().ff_cross_region_blend(_ptr(hs), _ptr(bitmasks), bitmasks.shape[1], _ptr(region_ids), B // 4, S, Cc,
                                           _DT[hs.dtype], _stream())
    _lib.check(rc, "ff_cross_region_blend")
    _count("ff_cross_region_blend")
    return hs


# ---------------------------------------------------------------------------------------------------------------
# UNet-body glue in channels-last layout (csrc/unet_glue.cu).  bf16 only, no fallback.
# ---------------------------------------------------------------------------------------------------------------
def _nhwc_view(x: torch.Tensor, name: str):
    """x: logical [N,C,H,W] with channels_last strides (or C == 1 / H*W == 1 ambiguity resolved by a dense NHWC
    permutation) -> (N, HW, C); raises when the memory is not dense NHWC."""
    if not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (freefine_b200 has no CPU path)")
    if x.dtype != torch.bfloat16:
        raise TypeError(f"{name} must be bfloat16, got {x.dtype}")
    if x.dim() != 4:
        raise ValueError(f"{name} must be [N,C,H,W], got {tuple(x.shape)}")
    if not x.permute(0, 2, 3, 1).is_contiguous():
        raise ValueError(f"{name} must be dense channels_last (NHWC) memory")
    N, Cc, H, W = x.shape
    return N, H * W, Cc


_GN_WS: dict = {}


def _gn_workspace(N: int, G: int, device) -> torch.Tensor:
    """Workspace of ff_group_norm_nhwc, one per (device, stream): consecutive calls on a stream are ordered, so the
    buffer is safely reused.  It carries the barrier counters of the single-read kernel: zero-initialised here, left
    zeroed by every call (include/freefine_b200.h)."""
    key = (device, _stream().value)
    need = int(_lib.load().ff_group_norm_ws_bytes(N, G))
    ws = _GN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _GN_WS[key] = ws
    return ws


def group_norm_nhwc(x, gamma, beta, groups, eps, add_nc=None, silu=False):
    """act(GroupNorm(x + add_nc[:, :, None, None])) on a channels_last bf16 [N,C,H,W] tensor -> same shape / strides.
    add_nc: fp32 [N,C] or None.  See ff_group_norm_nhwc."""
    N, HW, Cc = _nhwc_view(x, "x")
    _chk(gamma, torch.bfloat16, "gamma", 1)
    _chk(beta, torch.bfloat16, "beta", 1)
    if gamma.numel() != Cc or beta.numel() != Cc:
        raise ValueError("gamma / beta must have C elements")
    add_ld = Cc
    if add_nc is not None:      # fp32 [N,C]; rows may be a column block of a wider matrix (unit column stride)
        if not add_nc.is_cuda or add_nc.dtype != torch.float32 or add_nc.dim() != 2:
            raise TypeError("add_nc must be a 2-D float32 CUDA tensor")
        if tuple(add_nc.shape) != (N, Cc) or add_nc.stride(1) != 1 or (N > 1 and add_nc.stride(0) < Cc):
            raise ValueError(f"add_nc must be [{N},{Cc}] with unit column stride, got {tuple(add_nc.shape)} {add_nc.stride()}")
        add_ld = add_nc.stride(0) if N > 1 else Cc
    y = torch.empty_like(x)             # preserves the channels_last strides
    ws = _gn_workspace(N, groups, x.device)
    rc = _lib.load().ff_group_norm_nhwc(_ptr(x), _ptr(add_nc), add_ld, _ptr(gamma), _ptr(beta), _ptr(y), _ptr(ws), N, HW, Cc,
                                        groups, float(eps), int(bool(silu)), _stream())
    _lib.check(rc, "ff_group_norm_nhwc")
    _count("ff_group_norm_nhwc")
    return y


def bias_residual_nhwc(h, bias=None, res=None):
    """h + bias[None,:,None,None] + res on channels_last bf16 [N,C,H,W] tensors, written in place into h."""
    N, HW, Cc = _nhwc_view(h, "h")
    if bias is not None:
        _chk(bias, torch.bfloat16, "bias", 1)
        if bias.numel() != Cc:
            raise ValueError("bias must have C elements")
    if res is not None:
        if _nhwc_view(res, "res") != (N, HW, Cc):
            raise ValueError("res must have the shape of h")
    rc = _lib.load().ff_bias_residual_nhwc(_ptr(h), _ptr(bias), _ptr(res), _ptr(h), N * HW, Cc, _stream())
    _lib.check(rc, "ff_bias_residual_nhwc")
    _count("ff_bias_residual_nhwc")
    return h


def upsample2x_nhwc(x):
    """Nearest 2x up-sampling of a channels_last bf16 [N,C,H,W] tensor -> channels_last [N,C,2H,2W].  See ff_upsample2x_nhwc."""
    N, HW, Cc = _nhwc_view(x, "x")
    H, W = x.shape[2], x.shape[3]
    y = torch.empty((N, Cc, 2 * H, 2 * W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    rc = _lib.load().ff_upsample2x_nhwc(_ptr(x), _ptr(y), N, H, W, Cc, _stream())
    _lib.check(rc, "ff_upsample2x_nhwc")
    _count("ff_upsample2x_nhwc")
    return y


def concat_nhwc(a, b):
    """torch.cat([a, b], dim=1) of channels_last bf16 [N,C,H,W] tensors -> channels_last [N,Ca+Cb,H,W].  See ff_concat_nhwc."""
    N, HW, Ca = _nhwc_view(a, "a")
    Nb, HWb, Cb = _nhwc_view(b, "b")
    if (N, HW) != (Nb, HWb) or a.shape[2:] != b.shape[2:]:
        raise ValueError(f"concat_nhwc: shapes {tuple(a.shape)} and {tuple(b.shape)} differ outside the channel dimension")
    out = torch.empty((N, Ca + Cb, a.shape[2], a.shape[3]), dtype=a.dtype, device=a.device, memory_format=torch.channels_last)
    rc = _lib.load().ff_concat_nhwc(_ptr(a), _ptr(b), _ptr(out), N * HW, Ca, Cb, _stream())
    _lib.check(rc, "ff_concat_nhwc")
    _count("ff_concat_nhwc")
    return out


_LT_WS: dict = {}


def linear_bias_residual(x, weight, bias=None, res=None, out=None):
    """x [..., K] . weight[N, K]^T (+ bias[N]) (+ res [..., N]) -> bf16 [..., N] in ONE cuBLASLt GEMM (bias epilogue +
    beta * C): the `Linear(h) + hidden_states` tail of every transformer sub-block without its elementwise add.
    See ff_linear_bias_residual."""
    _chk(x, torch.bfloat16, "x")
    _chk(weight, torch.bfloat16, "weight", 2)
    N, K = weight.shape
    if x.shape[-1] != K:
        raise ValueError(f"x has {x.shape[-1]} features, weight expects {K}")
    M = x.numel() // K
    if bias is not None:
        _chk(bias, torch.bfloat16, "bias", 1)
        if bias.numel() != N:
            raise ValueError("bias must have N elements")
    oshape = tuple(x.shape[:-1]) + (N,)
    if res is not None:
        _chk(res, torch.bfloat16, "res")
        if tuple(res.shape) != oshape:
            raise ValueError(f"res must be {oshape}, got {tuple(res.shape)}")
    if out is None:
        out = torch.empty(oshape, dtype=torch.bfloat16, device=x.device)
    else:
        _chk(out, torch.bfloat16, "out")
        if tuple(out.shape) != oshape:
            raise ValueError(f"out must be {oshape}, got {tuple(out.shape)}")
    key = (x.device, _stream().value)
    ws = _LT_WS.get(key)
    if ws is None:
        ws = _LT_WS[key] = torch.empty(32 << 20, dtype=torch.uint8, device=x.device)
    rc = _lib.load().ff_linear_bias_residual(_ptr(x), _ptr(weight), _ptr(bias), _ptr(res), _ptr(out), M, N, K, _ptr(ws),
                                             ws.numel(), _stream())
    _lib.check(rc, "ff_linear_bias_residual")
    _count("ff_linear_bias_residual")
    return out


def geglu(h):
    """h bf16 [..., 2F] -> h[..., :F] * gelu(h[..., F:]) (erf gelu).  See ff_geglu."""
    _chk(h, torch.bfloat16, "h")
    F2 = h.shape[-1]
    if F2 % 2:
        raise ValueError("last dim must be even")
    M = h.numel() // F2
    out = torch.empty(*h.shape[:-1], F2 // 2, dtype=h.dtype, device=h.device)
    rc = _lib.load().ff_geglu(_ptr(h), _ptr(out), M, F2 // 2, _stream())
    _lib.check(rc, "ff_geglu")
    _count("ff_geglu")
    return out


def layer_norm(x, gamma, beta, eps):
    """LayerNorm over the last dim of a contiguous bf16 tensor.  See ff_layer_norm."""
    _chk(x, torch.bfloat16, "x")
    Cc = x.shape[-1]
    _chk(gamma, torch.bfloat16, "gamma", 1)
    _chk(beta, torch.bfloat16, "beta", 1)
    if gamma.numel() != Cc or beta.numel() != Cc:
        raise ValueError("gamma / beta must have C elements")
    y = torch.empty_like(x)
    rc = _lib.load().ff_layer_norm(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), x.numel() // Cc, Cc, float(eps), _stream())
    _lib.check(rc, "ff_layer_norm")
    _count("ff_layer_norm")
    return y


def to_device_bytes(arr: np.ndarray, device) -> torch.Tensor:
    """numpy structured array -> uint8 CUDA tensor (pinned staging, asynchronous copy on the current stream)."""
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1))
    if torch.cuda.is_available():
        t = t.pin_memory()
    return t.to(device, non_blocking=True)
