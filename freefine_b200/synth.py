"""Synthetic GeoBench-2D-shaped workload: there is no network for datasets, so benchmark edits are
seeded random images (smoothed so that they have structure), elliptical object masks and move/rotate/scale parameters
drawn like the reference driver's cases (evaluation/FreeFine/freefine_batch_infer_2d.py:44-54: the moved bounding box
must stay inside the image)."""
from __future__ import annotations

import numpy as np


def blob_mask(res: int, rng, frac=(0.08, 0.2)) -> np.ndarray:
    cy, cx = rng.uniform(0.3 * res, 0.7 * res, 2)
    ay, ax = rng.uniform(frac[0] * res, frac[1] * res, 2)
    yy, xx = np.mgrid[0:res, 0:res]
    return ((((yy - cy) / ay) ** 2 + ((xx - cx) / ax) ** 2) <= 1.0).astype(np.uint8)


def smooth_image(res: int, rng) -> np.ndarray:
    img = rng.integers(0, 256, (res // 8, res // 8, 3)).astype(np.float32)
    img = np.kron(img, np.ones((8, 8, 1), np.float32))                       # blocky low-frequency content
    k = 15
    ker = np.ones(k, np.float32) / k
    for ax in (0, 1):
        img = np.apply_along_axis(lambda v: np.convolve(np.pad(v, k // 2, mode="edge"), ker, mode="valid"), ax, img)
    return np.clip(img, 0, 255).astype(np.uint8)


def make_edit(index: int, res: int = 512):
    """Edit `index` of the synthetic sweep -> dict(image u8 [res,res,3], mask u8 0/1 [res,res], edit_param
    (dx,dy,rz,sx,sy), prompt)."""
    rng = np.random.default_rng(1000 + index)
    img = smooth_image(res, rng)
    m = blob_mask(res, rng)
    ys, xs = np.where(m)
    for _ in range(64):
        dx, dy = rng.uniform(-0.2 * res, 0.2 * res, 2)
        if xs.min() + dx >= 0 and xs.max() + dx < res and ys.min() + dy >= 0 and ys.max() + dy < res:
            break
    else:
        dx = dy = 0.0
    rz = float(rng.uniform(-30, 30))
    s = float(rng.uniform(0.7, 1.3))
    return dict(image=img, mask=m, edit_param=(float(dx), float(dy), rz, s, s), prompt="")


def make_batch(first: int, n: int, res: int = 512):
    es = [make_edit(first + i, res) for i in range(n)]
    return dict(images=np.stack([e["image"] for e in es]), masks=np.stack([e["mask"] for e in es]),
                edit_params=[e["edit_param"] for e in es], prompts=[e["prompt"] for e in es])
