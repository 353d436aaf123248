"""ctypes binding of libfreefine_b200.so (the C ABI declared in include/freefine_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.  The library is
built in-tree by `python -m freefine_b200.csrc.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FREEFINE_B200_LIB", os.path.join(_HERE, "lib", "libfreefine_b200.so"))

FF_MAX_PASS = 4
FF_DT_F32, FF_DT_BF16, FF_DT_F16 = 0, 1, 2
FF_PASS_KEY_INVERT, FF_PASS_ROW_XOR, FF_PASS_ROW_WEIGHT, FF_PASS_KEY2_INVERT = 1, 2, 4, 8
FF_PASS_KEY_PREFIX, FF_PASS_KEY2_PREFIX = 16, 32


class FFAttnPass(C.Structure):
    _fields_ = [("kv_stream", C.c_int32), ("key_mask", C.c_int32), ("row_mask", C.c_int32), ("flags", C.c_uint32),
                ("weight", C.c_float), ("kv_stream2", C.c_int32), ("key_mask2", C.c_int32), ("reserved", C.c_int32)]


class FFAttnHeadPlan(C.Structure):
    _fields_ = [("n_pass", C.c_int32), ("reserved", C.c_int32 * 3), ("passes", FFAttnPass * FF_MAX_PASS)]


class FFAttnArgs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p), ("plan", C.c_void_p),
                ("bitmasks", C.c_void_p), ("mask_popcount", C.c_void_p),
                ("n_streams", C.c_int32), ("n_kv_streams", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
                ("s_q", C.c_int32), ("s_kv", C.c_int32), ("n_masks", C.c_int32), ("mask_words", C.c_int32),
                ("out_dtype", C.c_int32), ("scale", C.c_float), ("v_dtype", C.c_int32), ("v_head_stride", C.c_int32)]


PLAN_BYTES = C.sizeof(FFAttnHeadPlan)      # 144
PASS_BYTES = C.sizeof(FFAttnPass)          # 32

# name -> (restype, argtypes); every symbol include/freefine_b200.h declares
SIGNATURES = {
    "ff_version": (C.c_int, []),
    "ff_last_error": (C.c_char_p, []),
    "ff_attn_masked_kv": (C.c_int, [C.POINTER(FFAttnArgs), C.c_void_p]),
    "ff_attn_v_head_stride": (C.c_int, [C.c_int32]),
    "ff_attn_plain_smallkv": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 5 + [C.c_float, C.c_int32, C.c_void_p]),
    "ff_debug_set_timeline": (C.c_int, [C.c_void_p]),
    "ff_kv_gather_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "ff_mask_downsample_pack": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ff_upsample2x_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ff_concat_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "ff_linear_bias_residual": (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "ff_mask_prep": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 7 + [C.c_void_p] * 6),
    "ff_dilate_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ff_warp_affine_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_void_p]),
    "ff_ddim_cfg_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ff_debug_set_trace": (C.c_int, [C.c_void_p]),
    "ff_ddim_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                               C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p]),
    "ff_ddim_cfg_step_compose": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ff_ddim_inv_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "ff_cross_region_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_int32, C.c_void_p]),
    "ff_group_norm_ws_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "ff_group_norm_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "ff_bias_residual_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                        C.c_void_p]),
    "ff_geglu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "ff_layer_norm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the library once; raises if it has not been built (no CPU / PyTorch fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m freefine_b200.csrc.build` "
                               "(freefine_b200 has no fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the header and the library disagree
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ff_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
