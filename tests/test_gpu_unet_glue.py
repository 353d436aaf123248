"""sm_100a UNet-body glue kernels (csrc/unet_glue.cu) through the C ABI against torch restatements of their documented
semantics (the same restatements that pin the host-side dataflow on CPU, tests/test_unet_fast_cpu.py), and the
channels-last fast path of the stand-in UNet against its plain PyTorch forward.
Tolerances: outputs are bf16 (half-ulp 2^-9 = 2e-3 relative) computed from fp32 statistics: rtol 8e-3 (two ulps), atol
2e-3; the UNet fast path must be at least as close to the fp32 forward as the eager bf16 forward is."""
import pytest
import torch
import torch.nn.functional as F

from test_unet_fast_cpu import ref_bias_residual_nhwc, ref_geglu, ref_group_norm_nhwc, ref_layer_norm

pytestmark = pytest.mark.gpu

RTOL, ATOL = 8e-3, 2e-3


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from freefine_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _nhwc(n, c, h, w, dev, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(n, c, h, w, generator=g) * scale + shift).to(dev).bfloat16()
    return x.contiguous(memory_format=torch.channels_last)


GN_SHAPES = [
    # N, C, H, W, G
    (2, 80, 8, 8, 16),        # tiny preset: 5 channels per group (vectors straddle groups)
    (3, 240, 16, 16, 16),     # 15 per group
    (2, 320, 64, 64, 32),     # SD1.5 64x64
    (2, 2560, 8, 8, 32),      # up-block concat: more channel vectors than threads
    (1, 960, 64, 64, 32),     # 30 per group
    (2, 640, 24, 24, 32),     # ragged pixel chunks (768^2 mid resolution)
    (1, 320, 96, 96, 32),     # 768^2: chunk count capped
    (5, 1280, 1, 1, 32),      # single pixel
    # register-resident small-image kernel at the SD1.5 sizes (bundles of 1 / 2 / 4 groups, 2..24 vectors per thread)
    (2, 1280, 8, 8, 32), (2, 1280, 16, 16, 32), (2, 2560, 16, 16, 32), (2, 1920, 16, 16, 32), (2, 640, 32, 32, 32),
    (1, 1280, 32, 32, 32), (2, 320, 32, 32, 32), (2, 1280, 24, 24, 32),
    (2, 960, 32, 32, 32),     # 61 vectors per thread in one CTA: two-kernel form (cluster form with FF_GN_CLUSTER=1)
    # ragged pixel splits (odd image sizes; last CTA of the cluster shorter in the cluster form)
    (1, 320, 72, 72, 32), (1, 640, 50, 50, 32), (1, 320, 47, 47, 32), (2, 1920, 32, 32, 32),
]


@pytest.mark.parametrize("shape", GN_SHAPES)
@pytest.mark.parametrize("silu,with_add", [(False, False), (True, True), (True, False)])
def test_group_norm_nhwc(dev, shape, silu, with_add):
    from freefine_b200 import ops
    n, c, h, w, G = shape
    x = _nhwc(n, c, h, w, dev, 11, scale=1.7, shift=0.4)
    g = torch.Generator(device="cpu").manual_seed(5)
    gamma = (1 + 0.3 * torch.randn(c, generator=g)).to(dev).bfloat16()
    beta = (0.2 * torch.randn(c, generator=g)).to(dev).bfloat16()
    add = (0.8 * torch.randn(n, c, generator=g)).to(dev) if with_add else None
    x0 = x.clone()
    got = ops.group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add, silu=silu)
    want = ref_group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add, silu=silu)
    assert torch.equal(x, x0)
    assert got.shape == x.shape and got.stride() == x.stride() and got.dtype == torch.bfloat16
    torch.testing.assert_close(got.float(), want.float(), rtol=RTOL, atol=ATOL)
    # bit-reproducible run to run (fixed summation order)
    assert torch.equal(got, ops.group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add, silu=silu))


def test_group_norm_cluster_switch(dev):
    """FF_GN_CLUSTER=1 (experiment switch, read once per process): the thread-block-cluster form of the register-resident
    kernel -- pixel split over 2 / 4 / 8 CTAs, statistics exchanged through distributed shared memory -- in a child process."""
    import os, subprocess, sys
    code = (
        "import sys, torch; sys.path.insert(0, 'tests')\n"
        "from test_unet_fast_cpu import ref_group_norm_nhwc\n"
        "from freefine_b200 import ops\n"
        "dev = torch.device('cuda:0'); g = torch.Generator(device='cpu').manual_seed(3)\n"
        "for (n, c, h, w) in ((2, 320, 64, 64), (1, 960, 64, 64), (2, 960, 32, 32), (1, 320, 47, 47), (1, 640, 50, 50), (1, 320, 96, 96)):\n"
        "    x = (torch.randn(n, c, h, w, generator=g) * 1.7 + 0.4).to(dev).bfloat16().contiguous(memory_format=torch.channels_last)\n"
        "    ga = (1 + 0.3 * torch.randn(c, generator=g)).to(dev).bfloat16(); be = (0.2 * torch.randn(c, generator=g)).to(dev).bfloat16()\n"
        "    add = (0.8 * torch.randn(n, c, generator=g)).to(dev)\n"
        "    got = ops.group_norm_nhwc(x, ga, be, 32, 1e-5, add_nc=add, silu=True)\n"
        "    want = ref_group_norm_nhwc(x, ga, be, 32, 1e-5, add_nc=add, silu=True)\n"
        "    torch.testing.assert_close(got.float(), want.float(), rtol=8e-3, atol=2e-3)\n"
        "    assert torch.equal(got, ops.group_norm_nhwc(x, ga, be, 32, 1e-5, add_nc=add, silu=True))\n"
        "print('cluster ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FF_GN_CLUSTER="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "cluster ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_group_norm_large_mean(dev):
    """|mean| >> std: the one-pass E[x^2] - E[x]^2 statistics stay inside tolerance at the offsets bf16 inputs allow."""
    from freefine_b200 import ops
    x = _nhwc(2, 320, 32, 32, dev, 3, scale=1.0, shift=30.0)
    w = torch.ones(320, device=dev).bfloat16()
    b = torch.zeros(320, device=dev).bfloat16()
    got = ops.group_norm_nhwc(x, w, b, 32, 1e-5)
    want = ref_group_norm_nhwc(x, w, b, 32, 1e-5)
    torch.testing.assert_close(got.float(), want.float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("shape", [(2, 80, 8, 8), (3, 320, 64, 64), (2, 1280, 16, 16), (1, 2560, 8, 8)])
@pytest.mark.parametrize("with_bias,with_res", [(True, True), (True, False), (False, True)])
def test_bias_residual_nhwc(dev, shape, with_bias, with_res):
    from freefine_b200 import ops
    n, c, h, w = shape
    hh = _nhwc(n, c, h, w, dev, 21)
    res = _nhwc(n, c, h, w, dev, 22) if with_res else None
    bias = torch.randn(c, device=dev).bfloat16() if with_bias else None
    want = ref_bias_residual_nhwc(hh.clone(), bias, res)
    got = ops.bias_residual_nhwc(hh, bias, res)
    assert got.data_ptr() == hh.data_ptr()
    torch.testing.assert_close(got.float(), want.float(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("M,Fdim", [(77, 320), (4096, 1280), (231, 640), (64, 5120), (3, 8), (5, 2000), (9, 248)])
def test_geglu(dev, M, Fdim):
    from freefine_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(31)
    h = (2.5 * torch.randn(2, M, 2 * Fdim, generator=g)).to(dev).bfloat16()
    got = ops.geglu(h)
    want = ref_geglu(h)
    assert got.shape == (2, M, Fdim)
    torch.testing.assert_close(got.float(), want.float(), rtol=RTOL, atol=ATOL)
    # the eager pair of kernels this replaces (same operation order and intermediate rounding)
    x, gate = h.chunk(2, dim=-1)
    eager = x * F.gelu(gate)
    torch.testing.assert_close(got.float(), eager.float(), rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("M,C", [(231, 80), (4096, 320), (1024, 640), (257, 1280), (16, 2048), (9, 8), (100, 160)])
def test_layer_norm(dev, M, C):
    from freefine_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(41)
    x = (1.3 * torch.randn(2, M, C, generator=g) + 0.5).to(dev).bfloat16()
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).to(dev).bfloat16()
    beta = (0.2 * torch.randn(C, generator=g)).to(dev).bfloat16()
    got = ops.layer_norm(x, gamma, beta, 1e-5)
    want = ref_layer_norm(x, gamma, beta, 1e-5)
    torch.testing.assert_close(got.float(), want.float(), rtol=RTOL, atol=ATOL)


def test_invalid_shapes_fail_loudly(dev):
    from freefine_b200 import ops
    x = _nhwc(2, 84, 8, 8, dev, 1)          # 84 channels: rows are not 16-byte multiples
    w = torch.ones(84, device=dev).bfloat16()
    with pytest.raises(RuntimeError):
        ops.group_norm_nhwc(x, w, w, 4, 1e-5)
    with pytest.raises(ValueError):
        ops.group_norm_nhwc(x.contiguous(), w, w, 4, 1e-5)      # NCHW memory
    with pytest.raises(TypeError):
        ops.layer_norm(torch.zeros(4, 32, device=dev), w, w, 1e-5)


@pytest.mark.parametrize("hw", [16, 32])
def test_unet_fast_path_matches_plain(dev, hw, monkeypatch):
    """bf16 channels-last fast path vs the fp32 plain forward of the same weights; the eager bf16 forward is the yard
    stick (the fused kernels round less often than the eager chain, so they must not be further away)."""
    from freefine_b200 import ops, standin
    g = torch.Generator(device="cpu").manual_seed(9)
    parts32 = standin.build_standin("tiny", device=dev)
    parts16 = standin.build_standin("tiny", device=dev, dtype=torch.bfloat16)
    x = torch.randn(4, 4, hw, hw, generator=g).to(dev)
    enc = torch.randn(4, 77, 64, generator=g).to(dev)
    t = torch.tensor(321, device=dev)
    with torch.no_grad():
        want = parts32.unet(x, t, enc)
        before = dict(ops.COUNTS)
        fast = parts16.unet(x.bfloat16(), t, enc.bfloat16())
        launched = {k: ops.COUNTS.get(k, 0) - before.get(k, 0) for k in
                    ("ff_group_norm_nhwc", "ff_bias_residual_nhwc", "ff_geglu", "ff_layer_norm")}
        monkeypatch.setattr(standin, "_fast", lambda x: False)
        eager = standin.build_standin("tiny", device=dev, dtype=torch.bfloat16).unet(x.bfloat16(), t, enc.bfloat16())
    assert fast.is_contiguous() and fast.shape == want.shape
    # 61 GroupNorms, 16 GEGLUs, 48 LayerNorms per UNet call went through the C ABI
    assert launched["ff_group_norm_nhwc"] == 61 and launched["ff_geglu"] == 16 and launched["ff_layer_norm"] == 48
    assert launched["ff_bias_residual_nhwc"] > 0
    rel = lambda a: float((a.float() - want).norm() / want.norm())
    e_fast, e_eager = rel(fast), rel(eager)
    print(f"rel-L2 vs fp32: fast {e_fast:.3e}  eager bf16 {e_eager:.3e}")
    assert e_fast < 3e-2 and e_fast <= 1.25 * e_eager + 1e-3, (e_fast, e_eager)


@pytest.mark.parametrize("M,N,K", [(4096, 320, 320), (1024, 640, 2560), (77, 1280, 1280), (131072, 320, 1280), (8, 64, 256)])
@pytest.mark.parametrize("with_bias,with_res", [(True, True), (False, True), (True, False)])
def test_linear_bias_residual(dev, M, N, K, with_bias, with_res):
    """ff_linear_bias_residual (one cuBLASLt GEMM with bias epilogue + beta * C) vs x @ W^T + b + res in fp32."""
    from freefine_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev).bfloat16()
    w = torch.randn(N, K, generator=g).to(dev).bfloat16()
    b = torch.randn(N, generator=g).to(dev).bfloat16() if with_bias else None
    r = torch.randn(M, N, generator=g).to(dev).bfloat16() if with_res else None
    r0 = r.clone() if with_res else None
    got = ops.linear_bias_residual(x, w, b, r)
    want = x.float() @ w.float().t()
    if with_bias:
        want = want + b.float()
    if with_res:
        want = want + r.float()
        assert torch.equal(r, r0)                     # the residual is read, not written, unless out aliases it
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    torch.testing.assert_close(got.float(), want, rtol=RTOL, atol=2e-2)
    assert torch.equal(got, ops.linear_bias_residual(x, w, b, r))          # repeatable
    if with_res:                                      # in-place form: out aliases res (cuBLASLt then rounds twice, like eager)
        got2 = ops.linear_bias_residual(x, w, b, r, out=r)
        assert got2.data_ptr() == r.data_ptr()
        torch.testing.assert_close(got2.float(), want, rtol=RTOL, atol=4e-2)


def test_fast_path_folds_residuals_into_the_projection_gemms(dev):
    """With the attention plugin registered (the bench's configuration) every `Linear(h) + hidden_states` of a transformer
    block -- both attention out-projections, the feed-forward out-projection, proj_out -- is ONE GEMM: 4 x 16 calls, no
    elementwise add; the result stays as close to the fp32 forward as the unfused fast path."""
    from freefine_b200 import ops, selfcheck
    pipe16, _ = selfcheck.build_pipeline(dev, torch.bfloat16)
    pipe32, _ = selfcheck.build_pipeline(dev, torch.float32)
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(4, 4, 16, 16, generator=g).to(dev)
    enc = torch.randn(4, 77, 64, generator=g).to(dev)
    t = torch.tensor(321, device=dev)
    with torch.no_grad():
        want = pipe32.unet(x, t, encoder_hidden_states=enc)
        pipe32.controller.reset()
        before = ops.COUNTS.get("ff_linear_bias_residual", 0)
        fast = pipe16.unet(x.bfloat16(), t, encoder_hidden_states=enc.bfloat16())
        pipe16.controller.reset()
    assert ops.COUNTS.get("ff_linear_bias_residual", 0) - before == 64
    rel = float((fast.float() - want).norm() / want.norm())
    print(f"rel-L2 vs fp32 with fused residual GEMMs: {rel:.3e}")
    assert rel < 3e-2, rel


@pytest.mark.parametrize("shape", [(4, 320, 64, 64, 32), (3, 1280, 16, 16, 32), (2, 640, 32, 32, 32), (2, 1280, 8, 8, 32)])
def test_group_norm_addend_as_column_block(dev, shape):
    """add_nc handed over as a column block of a wider [N, sum C] matrix (row stride > C): every GroupNorm kernel form."""
    from freefine_b200 import ops
    n, c, h, w, G = shape
    x = _nhwc(n, c, h, w, dev, 13, scale=1.3, shift=-0.2)
    g = torch.Generator(device="cpu").manual_seed(6)
    gamma = (1 + 0.3 * torch.randn(c, generator=g)).to(dev).bfloat16()
    beta = (0.2 * torch.randn(c, generator=g)).to(dev).bfloat16()
    wide = (0.8 * torch.randn(n, 3 * c + 40, generator=g)).to(dev)
    add = wide[:, c + 8:2 * c + 8]
    assert not add.is_contiguous() or n == 1
    got = ops.group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add, silu=True)
    want = ref_group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add.contiguous(), silu=True)
    torch.testing.assert_close(got.float(), want.float(), rtol=RTOL, atol=ATOL)
    assert torch.equal(got, ops.group_norm_nhwc(x, gamma, beta, G, 1e-5, add_nc=add.contiguous(), silu=True))


@pytest.mark.parametrize("shape", [(2, 640, 32, 32), (3, 1280, 8, 8), (1, 320, 5, 7), (2, 8, 3, 3)])
def test_upsample2x_nhwc_bit_exact(dev, shape):
    from freefine_b200 import ops
    n, c, h, w = shape
    x = _nhwc(n, c, h, w, dev, 51)
    got = ops.upsample2x_nhwc(x)
    want = F.interpolate(x, scale_factor=2.0, mode="nearest")
    assert got.shape == (n, c, 2 * h, 2 * w) and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", [(2, 1280, 640, 32, 32), (3, 320, 320, 64, 64), (1, 8, 16, 3, 5), (2, 1280, 1280, 8, 8)])
def test_concat_nhwc_bit_exact(dev, shape):
    from freefine_b200 import ops
    n, ca, cb, h, w = shape
    a, b = _nhwc(n, ca, h, w, dev, 52), _nhwc(n, cb, h, w, dev, 53)
    got = ops.concat_nhwc(a, b)
    assert got.shape == (n, ca + cb, h, w) and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, torch.cat([a, b], dim=1))
    with pytest.raises(ValueError):
        ops.concat_nhwc(a, _nhwc(n, cb, h + 1, w, dev, 54))
