"""Coarse geometric edit = affine warp + resample + mask-guided blend on the GPU (ff_warp_affine_blend).

Mirrors the reference's call surface:
* `param2theta(param, w, h)` and `wrapAffine_tensor(tensor, theta, dsize, mode, ...)`  (src/utils/geo_utils.py:292-341):
  the tensor-warp pair; `wrapAffine_tensor` here is numerically the reference function (grid_sample semantics, zeros
  padding, align_corners=False) -- parity target 1e-5, nearest-mode indices bit-exact;
* `re_edit_2d(src_img, src_mask, edit_param, inp_cur)`  (src/utils/vis_utils.py:210-274): rotation about the mask's
  bounding-box centre by -rz, anisotropic scale, translation; bilinear warp of the image, nearest warp of the mask,
  `np.where(mask, warped, inp_cur)`.  The reference does this on the CPU with cv2.warpAffine; here the 2x3 matrix is
  built with the reference's formulas and the warp/blend is ONE kernel launch.  cv2 interpolates in 1/32-pixel fixed
  point, so the image matches cv2 to its quantisation (a few grey levels at edges) while the mask matches exactly when
  the pixel-centre-exact theta is used (`cv2_theta`, SURVEY.md quirk Q11: `param2theta` itself is exact only for pure
  translations).
* `dilate_mask(mask, k)` (vis_utils.py:340-347).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


def param2theta(param, w, h):
    """reference geo_utils.py:292-302: forward pixel-space 2x3 matrix -> normalised inverse theta (as published)."""
    param = np.concatenate([np.asarray(param, dtype=np.float64), np.array([[0, 0, 1]], dtype=np.float64)])
    inv = np.linalg.inv(param)
    theta = np.zeros([2, 3])
    theta[0, 0] = inv[0, 0]
    theta[0, 1] = inv[0, 1] * h / w
    theta[0, 2] = inv[0, 2] * 2 / w + theta[0, 0] + theta[0, 1] - 1
    theta[1, 0] = inv[1, 0] * w / h
    theta[1, 1] = inv[1, 1]
    theta[1, 2] = inv[1, 2] * 2 / h + theta[1, 0] + theta[1, 1] - 1
    return theta


def cv2_theta(param, w, h):
    """Pixel-centre-exact theta for align_corners=False: param2theta plus the (1-a-b)/W, (1-d-e)/H offsets that make
    the normalised map agree with cv2.warpAffine's pixel map for rotations and scales too (quirk Q11)."""
    m = np.concatenate([np.asarray(param, dtype=np.float64), np.array([[0, 0, 1]], dtype=np.float64)])
    inv = np.linalg.inv(m)
    theta = param2theta(param, w, h)
    theta[0, 2] += (1 - inv[0, 0] - inv[0, 1]) / w
    theta[1, 2] += (1 - inv[1, 1] - inv[1, 0]) / h
    return theta


def wrapAffine_tensor(tensor, theta, dsize, mode='bilinear', padding_mode='zeros', align_corners=False, border_value=0):
    """reference geo_utils.py:304-341.  tensor [H,W] / [C,H,W] / [N,C,H,W] (CUDA), theta [2,3] or [N,2,3],
    dsize (width, height)."""
    if padding_mode != 'zeros' or align_corners:
        raise NotImplementedError("only padding_mode='zeros', align_corners=False (what the reference uses)")
    t = tensor
    while t.dim() < 4:
        t = t[None]
    theta = torch.as_tensor(theta, dtype=torch.float32)
    out = ops.warp_affine_blend(t.contiguous(), theta, dsize, mode=mode)
    if border_value != 0:
        ones = torch.ones((t.shape[0], 1) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
        inside = ops.warp_affine_blend(ones, theta, dsize, mode=mode)
        out = out + (1 - inside) * border_value
    return out


def edit_matrix(src_mask, edit_param):
    """The 2x3 forward matrix of re_edit_2d (vis_utils.py:220-250): cv2.getRotationMatrix2D about the mask's bbox
    centre by -rz (scale 1), then translation (+ the scale re-centring term) and per-axis scaling of the diagonal."""
    if src_mask.ndim == 3:
        src_mask = src_mask[:, :, 0]
    dx, dy, rz, sx, sy = edit_param
    ys, xs = np.where(src_mask)
    if len(ys) == 0:
        raise ValueError("re_edit_2d: empty source mask")
    cx, cy = (xs.max() + xs.min()) / 2, (ys.max() + ys.min()) / 2
    ang = math.radians(-rz)
    a, b = math.cos(ang), math.sin(ang)
    M = np.array([[a, b, (1 - a) * cx - b * cy], [-b, a, b * cx + (1 - a) * cy]], dtype=np.float64)
    M[0, 2] += dx + (1 - sx) * cx
    M[1, 2] += dy + (1 - sy) * cy
    M[0, 0] *= sx
    M[1, 1] *= sy
    return M


def re_edit_2d_device(imgs, masks, thetas, bgs):
    """Batched device form: imgs/bgs f32 [E,3,H,W], masks u8 [E,H,W], thetas f32 [E,2,3] ->
    (blended f32 [E,3,H,W], warped mask u8 0/1 [E,H,W]).  One kernel launch for the whole batch."""
    return ops.warp_affine_blend(imgs, thetas, mask_src=masks, bg=bgs, want_mask=True)


def re_edit_2d(src_img, src_mask, edit_param, inp_cur, device="cuda"):
    """reference vis_utils.py:210-274 -> (final_image u8 HWC, transformed_mask u8 0/255, trans_hole_image u8 HWC)."""
    if src_mask.ndim == 3:
        src_mask = src_mask[:, :, 0]
    H, W = src_mask.shape[:2]
    M = edit_matrix(src_mask, edit_param)
    theta = torch.tensor(cv2_theta(M, W, H), dtype=torch.float32)[None]
    img = torch.from_numpy(np.ascontiguousarray(src_img)).to(device).permute(2, 0, 1)[None].float().contiguous()
    msk = torch.from_numpy((src_mask != 0).astype(np.uint8)).to(device)[None].contiguous()
    hole = torch.where(msk[:, None] != 0, torch.zeros_like(img), img)
    bgs = torch.cat([torch.from_numpy(np.ascontiguousarray(inp_cur)).to(device).permute(2, 0, 1)[None].float(), hole])
    out, wm = ops.warp_affine_blend(img.expand(2, -1, -1, -1).contiguous(), theta.expand(2, 2, 3), mask_src=msk.expand(2, -1, -1).contiguous(),
                                    bg=bgs.contiguous(), want_mask=True)
    to_u8 = lambda t: t.round().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
    return to_u8(out[0]), (wm[0] * 255).cpu().numpy(), to_u8(out[1])


def re_edit_3d(src_img, src_mask, edit_param, inp_cur, ori_img_a, ori_mask_a, device="cuda"):
    """reference vis_utils.py:275-339: the same affine coarse edit as re_edit_2d applied to the re-oriented object
    (`src_img`/`src_mask` come from the depth / novel-view pre-step, out of scope here); the hole image is built from the
    ORIGINAL image and mask (`ori_img_a` with `ori_mask_a` blanked, :316).  ori_mask_a: [H,W] or [H,W,1|3] (numpy
    broadcasting against the HWC image, as in the reference).
    -> (final_image u8 HWC, transformed_mask u8 0/255, trans_hole_image u8 HWC)."""
    if src_mask.ndim == 3:
        src_mask = src_mask[:, :, 0]
    H, W = src_mask.shape[:2]
    M = edit_matrix(src_mask, edit_param)
    theta = torch.tensor(cv2_theta(M, W, H), dtype=torch.float32)[None]
    img = torch.from_numpy(np.ascontiguousarray(src_img)).to(device).permute(2, 0, 1)[None].float().contiguous()
    msk = torch.from_numpy((src_mask != 0).astype(np.uint8)).to(device)[None].contiguous()
    om = np.asarray(ori_mask_a)
    om = om[:, :, None] if om.ndim == 2 else om
    hole_np = np.where(om, 0, ori_img_a)                                         # :316
    hole = torch.from_numpy(np.ascontiguousarray(hole_np)).to(device).permute(2, 0, 1)[None].float()
    bgs = torch.cat([torch.from_numpy(np.ascontiguousarray(inp_cur)).to(device).permute(2, 0, 1)[None].float(), hole])
    out, wm = ops.warp_affine_blend(img.expand(2, -1, -1, -1).contiguous(), theta.expand(2, 2, 3),
                                    mask_src=msk.expand(2, -1, -1).contiguous(), bg=bgs.contiguous(), want_mask=True)
    to_u8 = lambda t: t.round().clamp(0, 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
    return to_u8(out[0]), (wm[0] * 255).cpu().numpy(), to_u8(out[1])


def dilate_mask(mask, dilate_factor=15, device="cuda"):
    """reference vis_utils.py:340-347 (cv2.dilate, k x k ones, anchor k//2, outside = 0) as a device max-filter."""
    k = int(dilate_factor)
    a = k // 2
    m = torch.from_numpy(np.ascontiguousarray(mask.astype(np.uint8))).to(device)
    x = (m[None, None] if m.dim() == 2 else m.permute(2, 0, 1)[None]).float()
    y = F.max_pool2d(F.pad(x, (a, k - 1 - a, a, k - 1 - a), value=0.0), kernel_size=k, stride=1)
    y = y[0, 0] if m.dim() == 2 else y[0].permute(1, 2, 0)
    return y.to(torch.uint8).cpu().numpy()
