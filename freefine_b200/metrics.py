"""GeoBench metrics on the outputs of the hot path: WRAP_E and the geometric part of MD.

* `calculate_we` = evaluation/metrics/wrap_error.py:5-21 -- mean absolute difference between the coarse input and the
  generated image inside the target mask, averaged over samples.
* `get_transform_coordinates` = evaluation/metrics/MD/mean_distance.py:84-111 -- where every source pixel lands under the
  annotated edit (translation / rotation about the mask's centre of mass / isotropic scale / a 3-D correspondence file).
* `mean_distance_from_features` = the inner loop of calculate_md (:146-166): for each key point the arg-max of the cosine
  similarity between its source feature and the edited image's feature map, and its distance to the transformed
  coordinate.  The DIFT feature extractor itself (SD-2.1 UNet features, MD/dift_sd.py) and SIFT key-point matching are
  outside the path (third-party models); this function takes the two feature maps and the key points.
"""
from __future__ import annotations

import numpy as np
import torch


def calculate_we(data: dict, image_label: str, read=None) -> float:
    """data: {img: {"instances": {ins: {edit: {coarse_input_path, <image_label>, tgt_mask_path}}}}}.  `read(path)` returns
    the image as an array in [0,255] (default: PIL, like the reference)."""
    if read is None:
        from PIL import Image
        read = lambda p: np.array(Image.open(p))
    wrap_e, num = 0.0, 0
    for image in data.values():
        for instance in image["instances"].values():
            for sample in instance.values():
                coarse, gen, tgt = (read(sample[k]) / 255 for k in ("coarse_input_path", image_label, "tgt_mask_path"))
                mask = np.repeat(tgt[..., np.newaxis], 3, axis=2)
                wrap_e += np.sum(np.abs(coarse * mask - gen * mask)) / mask.sum()
                num += 1
    return wrap_e / num


def wrap_error_batch(coarse: torch.Tensor, gen: torch.Tensor, tgt_mask: torch.Tensor) -> torch.Tensor:
    """The same per-sample quantity for tensors that are still on the device: coarse / gen uint8 [E,H,W,3], tgt_mask
    uint8 [E,H,W] in {0,255} -> float64 [E] (sum of |coarse - gen| / 255 over the mask, divided by 3 * mask.sum() / 255)."""
    m = tgt_mask.to(torch.float64) / 255
    d = (coarse.to(torch.float64) - gen.to(torch.float64)).abs() / 255
    return (d * m[..., None]).sum(dim=(1, 2, 3)) / (3 * m.sum(dim=(1, 2)))


def get_transform_coordinates(edit_param, size, mask, path_3D=None) -> np.ndarray:
    """[H,W,2] (row, col) target coordinate of every source pixel; edit_param = [dx,dy,dz,rx,ry,rz,sx,sy,sz]."""
    import cv2
    from scipy.ndimage import center_of_mass
    H, W = size
    if edit_param[0] != 0 or edit_param[1] != 0:
        ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        return np.stack((ii + edit_param[1], jj + edit_param[0]), axis=-1).astype(np.float64)
    if edit_param[5] != 0 or edit_param[6] != 1:
        center = center_of_mass(mask)
        if edit_param[5] != 0:
            matrix = cv2.getRotationMatrix2D(center, edit_param[5], scale=1.0)
        else:
            assert edit_param[6] == edit_param[7]
            sc = edit_param[6]
            matrix = np.array([[sc, 0, (1 - sc) * center[0]], [0, sc, (1 - sc) * center[1]]])
        y, x = np.meshgrid(np.arange(W), np.arange(H))
        pts = np.stack((x, y, np.ones_like(x)), axis=-1).reshape(-1, 3)
        return np.dot(pts, matrix.T).reshape(H, W, 2)
    return np.load(path_3D)[..., ::-1].copy()


def mean_distance_from_features(ft_source: torch.Tensor, ft_edited: torch.Tensor, kps, t_coords, max_points: int = 30):
    """ft_* [1,C,H,W] (already interpolated to the image size), kps [K,2] (row, col) -> list of K distances (float).
    One normalised [K,C] x [C,H*W] product instead of K cosine-similarity maps."""
    kps = np.asarray(kps)[:max_points]
    if len(kps) == 0:
        return []
    _, C, H, W = ft_edited.shape
    rows = torch.as_tensor(kps[:, 0], device=ft_source.device, dtype=torch.long)
    cols = torch.as_tensor(kps[:, 1], device=ft_source.device, dtype=torch.long)
    src = ft_source[0, :, rows, cols].t().float()                                   # [K,C]
    tgt = ft_edited[0].reshape(C, H * W).float()
    eps = 1e-8                                                                      # torch.nn.CosineSimilarity default
    sim = (src / src.norm(dim=1, keepdim=True).clamp_min(eps)) @ (tgt / tgt.norm(dim=0, keepdim=True).clamp_min(eps))
    best = sim.argmax(dim=1).cpu().numpy()                                          # first maximum, like np.argmax
    rc = np.stack(np.unravel_index(best, (H, W)), axis=1)
    tp = np.asarray(t_coords)[kps[:, 0], kps[:, 1]]
    # (tp - max_rc) in float64, cast to float32, then the norm -- the order of the reference (:163-165)
    return [float(np.linalg.norm((tp[i].astype(np.float64) - rc[i]).astype(np.float32))) for i in range(len(kps))]
