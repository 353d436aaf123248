"""GPU diagnostic (not a test): ONE small launch of every kernel behind the C ABI, for compute-sanitizer
(profiles/scripts/sanitize.sh).  Sizes are tiny so that memcheck / racecheck finish in seconds; every attention
instantiation is visited: legacy kernel (d = 8, 40 fp16-P; d = 40, 160 bf16x2), ring kernel (d = 80), masked TCA plans in
bit-vector and prefix mode, ragged K/V (77 keys)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from freefine_b200 import ops, plans

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def attn(S, d, heads=2, Skv=None, p_operand="f16", prefix=False):
    Skv = S if Skv is None else Skv
    q = torch.randn(4, S, heads * d, generator=g).to(dev).bfloat16()
    k = torch.randn(4, Skv, heads * d, generator=g).to(dev).bfloat16()
    v = torch.randn(4, Skv, heads * d, generator=g).to(dev).bfloat16()
    if Skv != S:
        plan, bits, pop, idx = plans.plain_plan(4, heads), None, None, None
    else:
        src = (torch.rand(S, generator=g) < 0.4).to(torch.uint8)
        tgt = (torch.rand(S, generator=g) < 0.3).to(torch.uint8)
        words = ops.mask_words(S)
        arr = np.zeros((2, words), np.uint32)
        for i, m in enumerate((src, tgt)):
            for j in np.nonzero(m.numpy())[0]:
                arr[i, j // 32] |= np.uint32(1) << np.uint32(j % 32)
        bits = torch.from_numpy(arr.view(np.int32)).to(dev)
        pop = torch.tensor([int(src.sum()), int(tgt.sum())], dtype=torch.int32, device=dev)
        plan = plans.tca_plan(1, heads, "tca", 0.5, lambda e: 0, lambda e: 1, prefix=prefix)
        idx = plans.kv_sort_index(torch.stack([src, tgt]).bool(), [-1, 0, -1, 0]).to(dev) if prefix else None
    k2, v2 = ops.kv_gather_cast(k, v, heads, idx, p_operand=p_operand)
    out = ops.attn_masked_kv(q, k2, v2, ops.to_device_bytes(plan, dev), heads, d ** -0.5, bits, pop, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all()), (S, d, p_operand)
    print(f"attn S={S} Skv={Skv} d={d} {p_operand} prefix={prefix}: ok")


for S, d, kw in ((192, 8, {}), (192, 40, {}), (192, 40, dict(prefix=True)), (192, 80, {}), (192, 80, dict(prefix=True)),
                 (192, 40, dict(p_operand="bf16x2")), (128, 160, dict(p_operand="bf16x2")), (192, 40, dict(Skv=77)), (192, 80, dict(Skv=77)),
                 (128, 160, {})):
    attn(S, d, **kw)
E, h, w = 2, 16, 16
eps4 = torch.randn(E, 4, 4, h, w, device=dev)
x = torch.randn(E, 2, 4, h, w, device=dev)
noise = torch.randn(E, 2, 4, h, w, device=dev)
cm = torch.randint(0, 3, (E, h, w), device=dev, dtype=torch.uint8)
k = dict(sqrt_1m_at=0.6, sqrt_at=0.8, sqrt_ap=0.85, c_ddim=0.52, c_ddpm=0.5, sigma=0.14)
ops.ddim_cfg_step(eps4, x, noise, cm, cm, 7.5, **k)
ops.ddim_step(x, x, noise, cm, **k)
ops.ddim_cfg_step_compose(torch.randn(E * 4, 4, h, w, device=dev), 4, x[:, 0].contiguous(), noise[:, 0].contiguous(), cm, cm, 7.5, **k)
ops.ddim_inv_step(x, x, 0.6, 0.8, 0.85, 0.52)
src = torch.randn(2, 6, 32, 32, device=dev)
th = torch.tensor([[[0.95, 0.2, 0.05], [-0.2, 0.95, -0.03]]], device=dev).expand(2, 2, 3).contiguous()
mask = (torch.rand(2, 32, 32, device=dev) > 0.5).to(torch.uint8)
ops.warp_affine_blend(src, th, mask_src=mask, bg=torch.randn_like(src), want_mask=True)
ops.warp_affine_blend(src, th, mode="nearest")
ops.mask_downsample_pack((torch.rand(3, 128, 128, device=dev) > 0.5).to(torch.uint8), 16, 16)
hs = torch.randn(4, 256, 64, device=dev)
bits, _ = ops.mask_downsample_pack((torch.rand(1, 128, 128, device=dev) > 0.5).to(torch.uint8), 16, 16)
ops.cross_region_blend(hs, bits, torch.zeros(1, dtype=torch.int32, device=dev))
xg = torch.randn(2, 64, 16, 16, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
ga, be = torch.ones(64, device=dev).bfloat16(), torch.zeros(64, device=dev).bfloat16()
ops.group_norm_nhwc(xg, ga, be, 16, 1e-5, add_nc=torch.randn(2, 64, device=dev), silu=True)
ops.bias_residual_nhwc(xg.clone(), ga, xg)
ops.geglu(torch.randn(2, 64, 256, device=dev).bfloat16())
ops.layer_norm(torch.randn(2, 64, 320, device=dev).bfloat16(), torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16(), 1e-5)
# round 2: the register-resident GroupNorm (bundles of 1 and 4 groups, ragged pixel count), the two-kernel form on a shape
# that does not fit it, the generic LayerNorm next to the sub-warp-row one above, mask preparation and dilation
for (n, c, h, w, G) in ((2, 1280, 8, 8, 32), (1, 320, 24, 24, 32), (1, 960, 32, 32, 32)):
    xs_ = torch.randn(n, c, h, w, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    ops.group_norm_nhwc(xs_, torch.ones(c, device=dev).bfloat16(), torch.zeros(c, device=dev).bfloat16(), G, 1e-5, silu=True)
ops.layer_norm(torch.randn(3, 17, 160, device=dev).bfloat16(), torch.ones(160, device=dev).bfloat16(), torch.zeros(160, device=dev).bfloat16(), 1e-5)
ops.layer_norm(torch.randn(1, 5, 1280, device=dev).bfloat16(), torch.ones(1280, device=dev).bfloat16(), torch.zeros(1280, device=dev).bfloat16(), 1e-5)
mm = lambda: ((torch.rand(2, 128, 128, device=dev) > 0.7).to(torch.uint8) * 255).contiguous()
ops.mask_prep(mm(), mm(), None, mm(), (16, 16), True, True)
ops.mask_prep(mm(), mm(), mm(), mm(), (16, 16), False, True)
ops.dilate_mask(mm(), 15)
xu = torch.randn(2, 64, 5, 7, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
ops.upsample2x_nhwc(xu)
ops.concat_nhwc(xu, torch.randn(2, 24, 5, 7, device=dev).bfloat16().contiguous(memory_format=torch.channels_last))
ops.linear_bias_residual(torch.randn(33, 64, device=dev).bfloat16(), torch.randn(40, 64, device=dev).bfloat16(),
                         torch.randn(40, device=dev).bfloat16(), torch.randn(33, 40, device=dev).bfloat16())
# round 2 (third session): short-key plain attention (ragged rows, 77 / 128 / 3 keys, two key blocks: 200 / 256 keys, every head
# width, both output types)
for (b_, hd_, d_, sq_, sk_, dt_) in ((2, 2, 40, 300, 77, torch.bfloat16), (1, 3, 80, 70, 128, torch.float32), (2, 2, 160, 17, 3, torch.bfloat16),
                                     (1, 2, 8, 33, 77, torch.float32), (1, 2, 160, 40, 200, torch.bfloat16), (1, 2, 40, 70, 256, torch.float32)):
    ops.attn_plain_smallkv(torch.randn(b_, sq_, hd_ * d_, device=dev).bfloat16(), torch.randn(b_, sk_, hd_ * d_, device=dev).bfloat16(),
                           torch.randn(b_, sk_, hd_ * d_, device=dev).bfloat16(), hd_, d_ ** -0.5, out_dtype=dt_)
torch.cuda.synchronize()
print("kernel tour ok:", dict(ops.COUNTS))
