"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 / numpy) of FreeFine's denoising hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this file.
The product (freefine_b200/) never does: it fails loudly when the CUDA extension is missing.

Every function cites the reference file:line (relative to the upstream repo root) that it restates.  The
restatement is written in *closed form per (stream, head)* (SURVEY.md section 8a), NOT as a transliteration of
the reference's materialised `[B*heads, S, S]` additive masks, so it doubles as the specification of what the
sm_100a kernels compute.

Pinning: tests/test_oracle_vs_reference.py compares every function here with the unmodified reference imported
via oracle/ref_import.py when /root/reference is present, and tests/test_oracle_golden.py compares it with the
committed fixtures in tests/golden/ (generated from the reference by oracle/make_golden.py) everywhere else.
The reference itself ships no tests / golden vectors for this path, and its third-party
UNet/VAE arithmetic (diffusers 0.18.0) is absent: at that boundary parity is pinned only on the shared stand-in.
"""
from __future__ import annotations

import math

import numpy as np
import torch

# ----------------------------------------------------------------------------------------------------------------
# mask / index logic  (integer work: must be bit-exact)
# ----------------------------------------------------------------------------------------------------------------

def d_ratio_for(h: int, w: int, seq: int) -> int:
    """attention.py:848  d_ratio = 2**int(log2(sqrt(h*w//seq)) + 0.5)"""
    return 2 ** int(math.log2((h * w // seq) ** 0.5) + 0.5)


def get_down_h_w(d_ratio: int, h: int, w: int, seq: int):
    """attention.py:713-733 (without the per-controller cache): ceil-halving from (h//8, w//8)."""
    r = d_ratio // 8
    nh, nw = h // 8, w // 8
    while r != 1:
        r //= 2
        nh = (nh + 1) // 2
        nw = (nw + 1) // 2
    assert nh * nw == seq, f"{nh}*{nw} != {seq}"
    return nh, nw


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """F.interpolate(mode='nearest') source index: floor(i * in/out) computed in fp32 like ATen
    (aten/src/ATen/native/UpSample.h nearest_neighbor_compute_source_index)."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def downsample_nearest(mask: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    iy = torch.from_numpy(nearest_index(oh, mask.shape[0]))
    ix = torch.from_numpy(nearest_index(ow, mask.shape[1]))
    return mask[iy][:, ix]


def process_mask_before_attention(mask: torch.Tensor, seq: int) -> torch.Tensor:
    """attention.py:841-855.  Returns the flattened [seq] mask at the layer's resolution, dtype preserved.
    Quirk kept: if mask.max() > 1 the mask is divided by its max and cast BACK to its dtype (uint8: floor)."""
    if mask.max() > 1:
        mask = (mask / mask.max()).to(mask.dtype)
    h, w = mask.shape
    dr = d_ratio_for(h, w, seq)
    ah, aw = get_down_h_w(dr, h, w, seq)
    return downsample_nearest(mask, ah, aw).flatten()


def pack_bits(flat01) -> np.ndarray:
    """bit i of word i//32 = (flat[i] != 0)   (the format ff_mask_downsample_pack writes)."""
    b = np.asarray(flat01).astype(bool)
    n = (b.size + 31) // 32
    pad = np.zeros(n * 32, dtype=bool)
    pad[: b.size] = b
    return (pad.reshape(n, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)


# ----------------------------------------------------------------------------------------------------------------
# attention  (fp32; tolerance 2e-3 max-abs for the bf16 tensor-core kernels)
# ----------------------------------------------------------------------------------------------------------------

def _split_heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    """attention.py:758-767 but kept 4-D: [B,S,C] -> [B,H,S,d]"""
    b, s, c = x.shape
    return x.reshape(b, s, heads, c // heads).permute(0, 2, 1, 3)


def _merge_heads(x: torch.Tensor) -> torch.Tensor:
    """attention.py:768-773: [B,H,S,d] -> [B,S,C]"""
    b, h, s, d = x.shape
    return x.permute(0, 2, 1, 3).reshape(b, s, h * d)


def _softmax_av(q, k, v, scale, allowed=None):
    """One head-pass: softmax_k(q.k*scale over allowed keys) @ v.  allowed: bool [Sq,Sk] or None.
    Q4 (attention.py:857): disallowed keys get finfo.min, so a row with NO allowed key is uniform over ALL keys."""
    s = (q @ k.transpose(-1, -2)) * scale
    if allowed is not None:
        none_allowed = ~allowed.any(dim=-1, keepdim=True)
        s = torch.where(allowed | none_allowed, s, torch.full_like(s, float("-inf")))
        s = torch.where(none_allowed.expand_as(s), torch.zeros_like(s), s)
    return torch.softmax(s, dim=-1) @ v


def plain_attention(q, k, v, heads, scale):
    """attention.py:395-404 / :1051-1058: unmasked attention on un-split [B,S,C] tensors."""
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    return _merge_heads(_softmax_av(qh, kh, vh, scale))


def q0_masked(heads: int, stream: int, head: int) -> bool:
    """Quirk Q0 (attention.py:859,881 vs :758-767): mask stack [M,1,M,1] is *tiled* over batch*heads, so
    batch*head index j = heads*stream + head receives stack entry j mod 4; entries 0 and 2 carry the region masks."""
    return ((heads * stream + head) % 4) in (0, 2)


def tca(q, k, v, heads, scale, src, tgt, method, cg, kind="edit"):
    """Temporal_contextal_attention (attention.py:1043-1091) and _bg (:1284-1324) in closed form.

    q,k,v: [4,S,C] streams [u_e,u_r,c_e,c_r];  src,tgt: [S] 0/1 masks at this layer's resolution
    (edit: src=fg_ref_mask keys, tgt=fg_retain_mask rows;  bg: src=tgt=obj mask, allowed keys = NOT obj for every row).
    """
    B, S, C = q.shape
    assert B == 4
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    srcb = torch.as_tensor(src).flatten().bool()
    tgtb = torch.as_tensor(tgt).flatten().bool()
    out = torch.empty_like(qh)
    kv_src = (1, 1, 3, 3)                      # attention.py:1033-1035
    for s in range(4):
        r = kv_src[s]
        for h in range(heads):
            if q0_masked(heads, s, h):
                if kind == "edit":             # rows in tgt read src keys, other rows read NOT-src keys (:1069/:1081)
                    allowed = torch.where(tgtb[:, None], srcb[None, :], ~srcb[None, :])
                else:                          # bg-gen: every row reads keys outside the object (:1310)
                    allowed = (~srcb)[None, :].expand(S, S)
            else:
                allowed = None
            o_ref = _softmax_av(qh[s, h], kh[r, h], vh[r, h], scale, allowed)
            if method == "mmsa":
                out[s, h] = o_ref
            else:
                o_self = _softmax_av(qh[s, h], kh[s, h], vh[s, h], scale)
                if kind == "edit":
                    out[s, h] = o_ref * cg + o_self * (1 - cg)        # :1083
                else:
                    out[s, h] = o_self * (1 - cg) + o_ref * cg        # :1316
    return _merge_heads(out)


def tca_compose(q, k, v, heads, scale, src_list, tgt_list, method, cg):
    """Temporal_contextal_attention_compose (attention.py:1092-1140): streams [u_e, r_1..r_N, c_e]; no Q0 quirk."""
    B, S, C = q.shape
    n = B - 2
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    self_h = _softmax_av(qh, kh, vh, scale)
    out = self_h.clone()
    for s in (0, B - 1):
        new = torch.zeros_like(qh[s])
        for i in range(n):
            srcb = torch.as_tensor(src_list[i]).flatten().bool()
            feat = torch.as_tensor(tgt_list[i]).flatten().to(q.dtype)
            allowed = srcb[None, :].expand(S, S)
            new = new + feat[None, :, None] * _softmax_av(qh[s], kh[1 + i], vh[1 + i], scale, allowed[None])
        out[s] = new if method == "mmsa" else new * cg + self_h[s] * (1 - cg)
    return _merge_heads(out)


def style_align(q, k, v, heads, scale, src=None, bg=False):
    """style_align_share_attention (attention.py:1142-1192): keys/values [self ; ref] under ONE softmax;
    'sdsa' (src given) masks the ref half by fg_ref_mask with the Q0 tiling (prepare_sdsa_mask :940-951).
    bg=True: style_align_share_attention_bg (:1193-1238) with prepare_sdsa_mask_for_bggen (:926-939): additive mask of
    1 - [ones ; obj], i.e. on the Q0-masked pairs the self half is masked out and the ref half admits keys outside obj."""
    B, S, C = q.shape
    qh, kh, vh = _split_heads(q, heads), _split_heads(k, heads), _split_heads(v, heads)
    out = torch.empty_like(qh)
    for s in range(B):
        r = 1 if s < B // 2 else B // 2 + 1
        for h in range(heads):
            kk = torch.cat([kh[s, h], kh[r, h]], 0)
            vv = torch.cat([vh[s, h], vh[r, h]], 0)
            allowed = None
            if src is not None and q0_masked(heads, s, h):
                srcb = torch.as_tensor(src).flatten().bool()
                if bg:
                    allowed = torch.cat([torch.zeros(S, dtype=torch.bool), ~srcb])[None, :].expand(S, 2 * S)
                else:
                    allowed = torch.cat([torch.ones(S, dtype=torch.bool), srcb])[None, :].expand(S, 2 * S)
            out[s, h] = _softmax_av(qh[s, h], kk, vv, scale, allowed)
    return _merge_heads(out)


def cross_local(q, k, v, heads, scale, region):
    """modulate_local_cross_attn (attention.py:1360-1393, _bg :1326-1357).  region: [S] at layer resolution,
    dtype preserved (uint8 `1-region` wraps exactly like the reference)."""
    hs = plain_attention(q, k, v, heads, scale)
    u_e, u_r, c_e, _ = hs
    region = torch.as_tensor(region).flatten()
    mod = region[:, None] * c_e + (1 - region)[:, None] * u_e
    return torch.stack([u_e, u_r, mod, u_r], dim=0)


def cross_local_compose(q, k, v, heads, scale, tgt_list, prompt_length):
    """modulate_local_cross_attn_compose (attention.py:1394-1432): q has nq streams, k/v have nq-1+prompt_length."""
    nq = q.shape[0]
    hu = plain_attention(q[: nq - 1], k[: nq - 1], v[: nq - 1], heads, scale)
    hc = torch.zeros_like(q[nq - 1:])
    for i in range(prompt_length):
        m = torch.as_tensor(tgt_list[i]).flatten()
        hc = hc + m[None, :, None] * plain_attention(q[nq - 1:], k[nq - 1 + i: nq + i], v[nq - 1 + i: nq + i], heads, scale)
    return torch.cat([hu, hc], 0)


# ----------------------------------------------------------------------------------------------------------------
# schedule / step math (fp32, reference operation order)
# ----------------------------------------------------------------------------------------------------------------

def make_alphas_cumprod() -> torch.Tensor:
    """SD1.5 scheduler: scaled-linear beta 0.00085->0.012, 1000 steps."""
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def timesteps_for(n: int) -> torch.Tensor:
    return (torch.arange(0, n) * (1000 // n)).flip(0).to(torch.int64) + 1


def linear_param(t, t1, t0, t2, end_scale=0.5):
    """model.py:438-455"""
    if t < t1 or t > t2:
        raise ValueError(f"t must be in [{t1}, {t2}]")
    if t <= t0:
        return 1 + (end_scale - 1) / (t0 - t1) * (t - t1)
    return end_scale + (-end_scale / (t2 - t0)) * (t - t0)


def inv_step(eps, t: int, x, alphas, n_steps: int):
    """model.py:109-132"""
    next_step = t
    tp = min(t - 1000 // n_steps, 999)
    a_t = alphas[tp] if tp >= 0 else alphas[0]
    a_next = alphas[next_step]
    beta = 1 - a_t
    pred_x0 = (x - beta ** 0.5 * eps) / a_t ** 0.5
    pred_dir = (1 - a_next) ** 0.5 * eps
    return a_next ** 0.5 * pred_x0 + pred_dir, pred_x0


def cfg_local(eps_u, eps_c, gs: float, cfg_mask=None):
    """model.py:605-611 (cfg_mask None -> :608 plain CFG)."""
    if cfg_mask is None:
        return eps_u + gs * (eps_c - eps_u)
    return eps_u + gs * (eps_c - eps_u) * cfg_mask


def ctrl_step(eps, t: int, x, mask, eta: float, noise, alphas, n_steps: int):
    """model.py:134-198 for the 2-stream [edit, ref] batch; mask [h,w] (uint8 arithmetic wraps: quirk Q1);
    noise = the randn_tensor(model_output.shape) draw of :186-188."""
    prev_t = t - 1000 // n_steps
    a_t = alphas[t]
    a_prev = alphas[prev_t] if prev_t > 0 else alphas[0]
    beta = 1 - a_t
    pred_x0 = (x - beta ** 0.5 * eps) / a_t ** 0.5
    a_prev_v = alphas[prev_t] if prev_t >= 0 else alphas[0]           # _get_variance :200-209 (>= 0)
    variance = ((1 - a_prev_v) / (1 - a_t)) * (1 - a_t / a_prev_v)
    std = eta * variance ** 0.5
    assert eps.shape[0] == 2
    std = torch.cat((std[None,], torch.zeros_like(std)[None,]))[:, None, None, None]
    m = mask.repeat(1, 4, 1, 1)
    m = torch.cat((m, torch.ones_like(m)))
    pdm = (1 - a_prev - std ** 2) ** 0.5 * eps * m
    pred_dir = (1 - a_prev) ** 0.5 * eps * (1 - m) + pdm
    x_prev = a_prev ** 0.5 * pred_x0 + pred_dir
    if eta > 0:
        x_prev = x_prev + std * noise * m
    return x_prev, pred_x0


# ----------------------------------------------------------------------------------------------------------------
# affine warp (grid_sample semantics, zeros padding, align_corners=False) + blend
# ----------------------------------------------------------------------------------------------------------------

def param2theta(param, w, h):
    """geo_utils.py:292-302"""
    param = np.concatenate([param, np.array([0, 0, 1], dtype=param.dtype)[None]])
    param = np.linalg.inv(param)
    theta = np.zeros([2, 3])
    theta[0, 0] = param[0, 0]
    theta[0, 1] = param[0, 1] * h / w
    theta[0, 2] = param[0, 2] * 2 / w + theta[0, 0] + theta[0, 1] - 1
    theta[1, 0] = param[1, 0] * w / h
    theta[1, 1] = param[1, 1]
    theta[1, 2] = param[1, 2] * 2 / h + theta[1, 0] + theta[1, 1] - 1
    return theta


def warp_affine(src: np.ndarray, theta: np.ndarray, dsize, mode="bilinear") -> np.ndarray:
    """wrapAffine_tensor (geo_utils.py:304-341) = F.affine_grid + F.grid_sample(padding 'zeros',
    align_corners=False), restated in fp32 numpy with ATen's operation order:
      base x_j = (2j+1)/W - 1 ; grid = base @ theta^T ; ix = ((gx+1)*W_in - 1)/2 ; bilinear with zero OOB taps.
    src [N,C,H,W] fp32, theta [2,3] fp32, dsize (width,height)."""
    src = np.asarray(src, np.float32)
    th = np.asarray(theta, np.float32)
    N, C, H, W = src.shape
    dw, dh = dsize
    xs = ((np.arange(dw, dtype=np.float32) * 2 + 1) / np.float32(dw) - 1).astype(np.float32)
    ys = ((np.arange(dh, dtype=np.float32) * 2 + 1) / np.float32(dh) - 1).astype(np.float32)
    gx = (xs[None, :] * th[0, 0] + ys[:, None] * th[0, 1]).astype(np.float32) + th[0, 2]
    gy = (xs[None, :] * th[1, 0] + ys[:, None] * th[1, 1]).astype(np.float32) + th[1, 2]
    ix = ((gx + 1) * np.float32(W) - 1) / 2
    iy = ((gy + 1) * np.float32(H) - 1) / 2
    out = np.zeros((N, C, dh, dw), np.float32)
    if mode == "nearest":
        xi = np.rint(ix).astype(np.int64)      # std::nearbyint: ties to even
        yi = np.rint(iy).astype(np.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        out[:, :, ok] = src[:, :, yi[ok], xi[ok]]
        return out
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wts = [((x1 - ix) * (y1 - iy), x0, y0), ((ix - x0) * (y1 - iy), x1, y0),
           ((x1 - ix) * (iy - y0), x0, y1), ((ix - x0) * (iy - y0), x1, y1)]
    for wgt, xx, yy in wts:
        xi, yi = xx.astype(np.int64), yy.astype(np.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        tap = np.zeros((N, C, dh, dw), np.float32)
        tap[:, :, ok] = src[:, :, yi[ok], xi[ok]]
        out = out + tap * wgt.astype(np.float32)[None, None]
    return out


def warp_blend(src, theta, mask_src, bg):
    """Mask-guided blend after the warp (vis_utils.py:252-256,272 in tensor form): the mask is warped with
    nearest sampling, out = where(warped_mask != 0, warped_src, bg)."""
    N, C, H, W = bg.shape
    ws = warp_affine(src, theta, (W, H), "bilinear")
    wm = warp_affine(np.asarray(mask_src, np.float32)[None, None], theta, (W, H), "nearest")[0, 0]
    return np.where(wm[None, None] != 0, ws, np.asarray(bg, np.float32)), (wm != 0).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------
# per-edit mask preparation (uint8 algebra incl. wrap-around quirk Q1)
# ----------------------------------------------------------------------------------------------------------------

def re_edit_2d(src_img: np.ndarray, src_mask: np.ndarray, edit_param, inp_cur: np.ndarray):
    """Coarse 2-D edit on the CPU exactly as the reference does it with OpenCV (src/utils/vis_utils.py:210-274): rotation by
    -rz about the centre of the mask's bounding box (cv2.getRotationMatrix2D, scale 1), translation (dx, dy) plus the
    re-centring term (1 - s) * centre, the diagonal scaled by (sx, sy); bilinear warp of the image, nearest warp of the
    mask, np.where blends.  Returns (final image, warped mask * 255, image with the hole)."""
    import cv2
    m = src_mask[:, :, 0] if src_mask.ndim == 3 else src_mask
    dx, dy, rz, sx, sy = edit_param
    h, w = m.shape[:2]
    ys, xs = np.where(m)
    cx, cy = (xs.max() + xs.min()) / 2, (ys.max() + ys.min()) / 2
    M = cv2.getRotationMatrix2D((cx, cy), -rz, 1)
    M[0, 2] += dx + (1 - sx) * cx
    M[1, 2] += dy + (1 - sy) * cy
    M[0, 0] *= sx
    M[1, 1] *= sy
    warped = cv2.warpAffine(src_img, M, (w, h))
    wmask = cv2.warpAffine(m.astype(np.uint8), M, (w, h), flags=cv2.INTER_NEAREST).astype(bool)
    hole = np.where(m[:, :, None].astype(bool), 0, src_img)
    return (np.where(wmask[:, :, None], warped, inp_cur), wmask.astype(np.uint8) * 255,
            np.where(wmask[:, :, None], warped, hole))


def dilate_mask(mask: np.ndarray, k: int) -> np.ndarray:
    """model.py:927-934 / vis_utils.py:340-347: cv2.dilate with a k x k ones kernel (anchor at k//2), restated as a
    sliding max with OpenCV's border handling for dilation (out-of-image = minimum)."""
    m = mask.astype(np.uint8)
    a = k // 2
    H, W = m.shape[:2]
    pad = np.zeros((H + k - 1, W + k - 1) + m.shape[2:], np.uint8)
    pad[a:a + H, a:a + W] = m
    out = np.zeros_like(m)
    for dy in range(k):
        for dx in range(k):
            np.maximum(out, pad[dy:dy + H, dx:dx + W], out=out)
    return out


def prepare_tensor_mask(mask: np.ndarray, w: int, h: int) -> torch.Tensor:
    """model.py:1622-1639 (binary=True)"""
    if mask.ndim == 3:
        mask = mask[:, :, 0]
    t = torch.tensor(mask)
    t = downsample_nearest(t, h, w).clone()
    t[t > 0.0] = 1.0
    return t


def prepare_various_mask(shifted_mask, ori_mask, draw_mask, w, h, lat_h, lat_w, use_auto_draw=False, cons_area=None,
                         reduce_inp_artifacts=False):
    """model.py:1432-1512 -> (fg_mask, shifted, ori, completion [lat], local_var [lat]); uint8 wraps kept."""
    P = lambda m: prepare_tensor_mask(m, w, h)
    if not use_auto_draw:
        if not reduce_inp_artifacts:
            sh, ori = P(shifted_mask), P(ori_mask)
            flex = P(draw_mask) * (1 - sh)
            fg = flex + sh
            fg[fg > 0] = 1.0
            comp, lvar = flex, flex
        else:
            dil = P(dilate_mask(ori_mask, 30))
            cons, sh, ori = P(cons_area), P(shifted_mask), P(ori_mask)
            flex = P(draw_mask) * (1 - sh)
            fg = flex + sh
            fg[fg > 0] = 1.0
            comp = flex
            lvar = (1 - cons) * (1 - sh) * dil + flex
            lvar[lvar > 0] = 1
    else:
        if not reduce_inp_artifacts:
            dil_t = P(dilate_mask(shifted_mask, 15))
            sh, ori, cons = P(shifted_mask), P(ori_mask), P(cons_area)
            fg = sh
            cons = cons - ori
            comp = (1 - cons) * (1 - sh) * dil_t
            lvar = comp
        else:
            dil_t = P(dilate_mask(shifted_mask, 15))
            dil = P(dilate_mask(ori_mask, 30))
            sh, ori, cons = P(shifted_mask), P(ori_mask), P(cons_area)
            fg = sh
            cons = cons - ori
            comp = dil + dil_t
            comp[comp > 0] = 1
            comp = comp * ((1 - cons) * (1 - sh))
            lvar = comp
    comp = downsample_nearest(comp, lat_h, lat_w)
    lvar = downsample_nearest(lvar, lat_h, lat_w)
    return fg, sh, ori, comp, lvar
