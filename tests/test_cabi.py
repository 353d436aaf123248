"""The C-ABI shared library loads and exports every symbol include/freefine_b200.h declares (no compute calls:
this runs without a GPU), the ctypes struct layouts match the header, and argument validation fails loudly."""
import ctypes as C
import os
import re

import pytest

from freefine_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from freefine_b200.csrc.build import build
        build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "freefine_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(ff_\w+)\s*\(", hdr, re.M))
    assert declared == set(_lib.SIGNATURES), (declared, set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ff_version() == 100


def test_struct_layout():
    assert C.sizeof(_lib.FFAttnPass) == 32
    assert C.sizeof(_lib.FFAttnHeadPlan) == 16 + 4 * 32
    assert _lib.FFAttnArgs.n_streams.offset == 56 and C.sizeof(_lib.FFAttnArgs) == 104
    from freefine_b200 import plans
    assert plans.PLAN_DTYPE.itemsize == C.sizeof(_lib.FFAttnHeadPlan)
    assert plans.PLAN_DTYPE.fields["passes"][1] == _lib.FFAttnHeadPlan.passes.offset
    for f in ("kv_stream", "key_mask", "row_mask", "flags", "weight", "kv_stream2", "key_mask2"):
        assert plans.PASS_DTYPE.fields[f][1] == getattr(_lib.FFAttnPass, f).offset, f


def test_invalid_arguments_fail_loudly(lib):
    # null pointers / bad shapes are rejected on the host before any CUDA call
    assert lib.ff_ddim_inv_step(None, None, 0.0, 1.0, 1.0, 0.0, None, None, 16, None) == -1
    assert b"null" in lib.ff_last_error()
    assert lib.ff_mask_downsample_pack(None, 1, 8, 8, 4, 4, None, 1, None, None) == -1
    assert lib.ff_attn_masked_kv(None, None) == -1
    a = _lib.FFAttnArgs()
    assert lib.ff_attn_masked_kv(C.byref(a), None) == -1
    with pytest.raises(RuntimeError):
        _lib.check(-1, "x")


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under freefine_b200/ may import it (no CPU fallback)."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "freefine_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), os.path.join(dp, f)


def test_v_head_stride_host_function(lib):
    """ff_attn_v_head_stride is pure host code: channels per head of the staged V (head_dim + ones column, padded to the
    N granularity of the PV tensor-core instruction)."""
    for d, want in ((8, 16), (16, 48), (40, 48), (48, 96), (80, 96), (88, 176), (160, 176)):
        assert lib.ff_attn_v_head_stride(d) == want, d
        assert want > d and want % 16 == 0
