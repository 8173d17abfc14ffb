"""Seeded inputs shared by the golden-vector generator (run against the reference in the authoring container)
and by the tests (run anywhere): only the reference's *outputs* are stored in the .npz files, the inputs are
regenerated from the seed.  numpy's PCG64 streams are stable across platforms.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
N_CASES = 40
OUT_SIZE = 129


def make_image(rng: np.random.Generator, w: int, h: int, kind: str) -> np.ndarray:
    """uint8 [h, w] source: 'noise' = iid uniform (worst case for interpolation parity), 'smooth' = natural-like."""
    if kind == "noise":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = (np.sin(x / 17.0) + np.cos(y / 23.0) + 2.0) / 4.0 * 255.0
    return np.clip(np.rint(base + rng.normal(0.0, 8.0, (h, w))), 0, 255).astype(np.uint8)


def make_labels(rng: np.random.Generator, w: int, h: int, roi_w=None, roi_h=None):
    """roi / coord / pose / pt3d_68 in the ranges of the bundled aflw2kmini.h5."""
    bw = rng.uniform(147, 250) if roi_w is None else roi_w
    bh = rng.uniform(147, 250) if roi_h is None else roi_h
    sx, sy = w / 450.0, h / 450.0
    bw, bh = bw * sx, bh * sy
    cx = w / 2 + rng.uniform(-40, 40) * sx
    cy = h / 2 + rng.uniform(-40, 40) * sy
    roi = np.asarray([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], F32)
    coord = np.asarray([cx, cy, 0.5 * max(bw, bh)], F32)
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    pts = np.empty((68, 3), F32)
    pts[:, 0] = rng.uniform(roi[0], roi[2], 68)
    pts[:, 1] = rng.uniform(roi[1], roi[3], 68)
    pts[:, 2] = rng.normal(0, 30, 68)
    shapeparam = rng.standard_normal(50).astype(F32)
    return dict(roi=roi, coord=coord, pose=q.astype(F32), pt3d_68=pts, shapeparam=shapeparam)


def make_case(i: int) -> dict:
    """Case i of the golden set: source image, labels and every augmentation draw."""
    rng = np.random.default_rng(1000 + i)
    kind = "noise" if i % 2 == 0 else "smooth"
    w, h = [(450, 450), (450, 450), (320, 240), (200, 180), (640, 480)][i % 5]
    roi_w = roi_h = None
    scale = F32(np.clip(0.1 * rng.standard_normal(), -0.5, 0.5) + 1.1)
    translation = np.clip(0.5 * rng.standard_normal(2), -1, 1).astype(F32)
    angle = F32(0.0)
    if i % 3 == 1:
        angle = F32(np.pi * 30.0 / 180.0) * F32(1 if i % 2 else -1)
    if i % 7 == 3:
        angle = F32(rng.uniform(-0.7, 0.7))  # arbitrary angle (cos/sin rounding caveat lives here)
    do_flip = bool(i % 2 == 1) if i % 4 else bool(rng.integers(0, 2))
    rot_dir = int([0, 0, 0, 0, 1, 0, 0, -1][i % 8])
    special = None
    if i == 30:  # exact 2x area path: 258-px square view box
        w, h, kind, special = 450, 450, "noise", "box258"
    if i == 31:  # tiny face: crop smaller than the output -> INTER_LINEAR up-scaling
        w, h, special = 200, 180, "tiny"
    if i == 32:  # tiny face, rotated -> direct up-scaling warp
        w, h, special, angle = 200, 180, "tiny", F32(np.pi * 30.0 / 180.0)
    if i == 33:  # box hanging far over the image border
        special = "border"
    lab = make_labels(rng, w, h, roi_w, roi_h)
    if special == "box258":
        lab["roi"] = np.asarray([100.0, 96.0, 358.0, 354.0], F32)
        scale, translation, angle = F32(1.0), np.zeros(2, F32), F32(0.0)
    if special == "tiny":
        cx, cy = lab["coord"][:2]
        lab["roi"] = np.asarray([cx - 35, cy - 40, cx + 35, cy + 40], F32)
    if special == "border":
        lab["roi"] = np.asarray([-60.0, 20.0, 150.0, 260.0], F32)
    img = make_image(rng, w, h, kind)
    return dict(index=i, wh=(w, h), image=img, scale=scale, angle=angle, translation=translation, do_flip=do_flip,
                rot_dir=rot_dir, out_size=OUT_SIZE, **lab)
