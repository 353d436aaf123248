"""Dataflow of the channels-last fast path of the stand-in UNet (freefine_b200/standin.py), checked WITHOUT a GPU: the
fused C-ABI ops are replaced -- in this test only -- by torch restatements of their documented semantics
(include/freefine_b200.h), the fast branch is forced on, and the result must equal the plain PyTorch forward.  This
pins the host-side restructuring (bias / time-embedding folding, skip connections, token views, memory formats); the
kernels themselves are checked against the same restatements on the GPU (tests/test_gpu_unet_glue.py)."""
import pytest
import torch
import torch.nn.functional as F

from freefine_b200 import ops, standin


def _dense_nhwc(x):
    assert x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous(), "activation is not dense NHWC"


def ref_group_norm_nhwc(x, gamma, beta, groups, eps, add_nc=None, silu=False):
    _dense_nhwc(x)
    v = x.float()
    if add_nc is not None:
        assert add_nc.dtype == torch.float32 and tuple(add_nc.shape) == (x.shape[0], x.shape[1])
        v = v + add_nc[:, :, None, None]
    y = F.group_norm(v, groups, gamma.float(), beta.float(), eps)
    if silu:
        y = F.silu(y)
    return y.to(x.dtype).contiguous(memory_format=torch.channels_last)


def ref_bias_residual_nhwc(h, bias=None, res=None):
    _dense_nhwc(h)
    v = h.float()
    if bias is not None:
        v = v + bias.float()[None, :, None, None]
    if res is not None:
        _dense_nhwc(res)
        v = v + res.float()
    h.copy_(v.to(h.dtype))
    return h


def ref_geglu(h):
    assert h.is_contiguous()
    x, gate = h.float().chunk(2, dim=-1)
    return (x * F.gelu(gate).to(h.dtype).float()).to(h.dtype)


def ref_layer_norm(x, gamma, beta, eps):
    assert x.is_contiguous()
    return F.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), eps).to(x.dtype)


def ref_linear_bias_residual(x, weight, bias=None, res=None, out=None):
    assert x.is_contiguous() and weight.is_contiguous() and (res is None or res.is_contiguous())
    v = x.float() @ weight.float().t()
    if bias is not None:
        v = v + bias.float()
    if res is not None:
        assert res.shape == v.shape
        v = v + res.float()
    return v.to(x.dtype)


def ref_upsample2x_nhwc(x):
    _dense_nhwc(x)
    return F.interpolate(x, scale_factor=2.0, mode="nearest").contiguous(memory_format=torch.channels_last)


def ref_concat_nhwc(a, b):
    _dense_nhwc(a)
    _dense_nhwc(b)
    return torch.cat([a, b], dim=1).contiguous(memory_format=torch.channels_last)


@pytest.fixture
def forced_fast(monkeypatch):
    monkeypatch.setattr(standin, "_fast", lambda x: True)
    monkeypatch.setattr(ops, "group_norm_nhwc", ref_group_norm_nhwc)
    monkeypatch.setattr(ops, "bias_residual_nhwc", ref_bias_residual_nhwc)
    monkeypatch.setattr(ops, "geglu", ref_geglu)
    monkeypatch.setattr(ops, "layer_norm", ref_layer_norm)
    monkeypatch.setattr(ops, "linear_bias_residual", ref_linear_bias_residual)
    monkeypatch.setattr(ops, "upsample2x_nhwc", ref_upsample2x_nhwc)
    monkeypatch.setattr(ops, "concat_nhwc", ref_concat_nhwc)


def _inputs(parts, n=3, hw=16):
    g = torch.Generator().manual_seed(3)
    cfg = parts.unet.config
    x = torch.randn(n, cfg.in_channels, hw, hw, generator=g)
    enc = torch.randn(n, 77, cfg.cross_attention_dim, generator=g)
    return x, torch.tensor(401), enc


def test_fast_path_equals_plain_forward(forced_fast, monkeypatch):
    parts = standin.build_standin("tiny")
    x, t, enc = _inputs(parts)
    with torch.no_grad():
        got = parts.unet(x, t, enc)
        monkeypatch.setattr(standin, "_fast", lambda x: False)
        want = standin.build_standin("tiny").unet(x, t, enc)
    assert got.is_contiguous() and got.shape == want.shape
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-5, err


def test_fast_path_with_registered_controller(forced_fast, monkeypatch):
    """The attention-processor hook (register_attention_control) sees the same [B,S,C] projections on the fast path:
    a recording processor must be called once per Attention module with identical shapes."""
    parts = standin.build_standin("tiny")
    x, t, enc = _inputs(parts, n=2, hw=8)
    seen = {}

    def hook(tag):
        def pre(mod, args, kwargs):
            seen.setdefault(tag, []).append(tuple(args[0].shape))
        return pre

    for m in parts.unet.modules():
        if m.__class__.__name__ == "Attention":
            m.register_forward_pre_hook(hook("fast"), with_kwargs=True)
    with torch.no_grad():
        parts.unet(x, t, enc)
        monkeypatch.setattr(standin, "_fast", lambda x: False)
        for m in parts.unet.modules():
            m._forward_pre_hooks.clear()
            if m.__class__.__name__ == "Attention":
                m.register_forward_pre_hook(hook("plain"), with_kwargs=True)
        parts.unet(x, t, enc)
    assert len(seen["fast"]) == 32 and seen["fast"] == seen["plain"]


def test_ops_reject_cpu_tensors():
    """The real ops have no CPU path: they raise before touching the library."""
    x = torch.zeros(2, 16, 4, 4, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.ones(16, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.group_norm_nhwc(x, w, w, 4, 1e-5)
    with pytest.raises(RuntimeError):
        ops.bias_residual_nhwc(x, w)
    with pytest.raises(RuntimeError):
        ops.geglu(torch.zeros(4, 32, dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):
        ops.layer_norm(torch.zeros(4, 32, dtype=torch.bfloat16), torch.ones(32, dtype=torch.bfloat16),
                       torch.ones(32, dtype=torch.bfloat16), 1e-5)
