"""Label updates under a 2-D similarity -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows trackertraincode/datatransformation/tensors/affinetrafo.py:37-148,
trackertraincode/neuralnets/torchquaternion.py:23-48,70-91 and trackertraincode/facemodel/keypoints68.py:7-76.
All arrays are float32; `tr` is [..., 2, 3] and broadcasts over leading sample dims.
"""
from __future__ import annotations

import numpy as np

from . import affine

F32 = np.float32

# Field categories, same string values as FieldCategory (trackertraincode/datasets/dshdf5pose.py:21-28).
CAT_GENERAL, CAT_IMAGE, CAT_QUAT, CAT_XYS, CAT_ROI, CAT_POINTS, CAT_SEMSEG = "", "img", "q", "xys", "roi", "pts", "seg"
IMAGELIKE = (CAT_IMAGE, CAT_SEMSEG)


def _build_flip_map():
    """Left/right landmark permutation of the 68-point scheme (keypoints68.py:7-76), built from its structure:
    jaw 0-16 and brows 17-26 reverse; nose bridge 27-30 fixed; nostrils 31-35 reverse; eyes swap with the
    upper/lower lid order kept; outer lips 48-54 / 55-59 and inner lips 60-64 / 65-67 reverse."""
    m = list(range(68))
    m[0:17] = range(16, -1, -1)
    m[17:27] = range(26, 16, -1)
    m[31:36] = range(35, 30, -1)
    m[36:42] = [45, 44, 43, 42, 47, 46]
    m[42:48] = [39, 38, 37, 36, 41, 40]
    m[48:55] = range(54, 47, -1)
    m[55:60] = range(59, 54, -1)
    m[60:65] = range(64, 59, -1)
    m[65:68] = range(67, 64, -1)
    return np.asarray(m, dtype=np.int64)


FLIP_MAP = _build_flip_map()


def transform_points(tr, pts):
    """affinetrafo.py:37-58.  xy' = R xy + t;  z' = sqrt(|det|) z (never mirrored)."""
    tr = np.asarray(tr, F32)
    pts = np.asarray(pts, F32)
    m = tr[..., None, :, :] if pts.ndim == tr.ndim else tr  # inject the point axis
    out = np.empty_like(pts)
    out[..., :2] = affine.affinevecmul(m, pts[..., :2])
    if pts.shape[-1] == 3:
        out[..., 2] = np.sqrt(np.abs(affine.det(tr)))[..., None] * pts[..., 2]
    return out


def transform_keypoints(tr, pts):
    """affinetrafo.py:61-72: transform, then swap left/right landmarks where the map mirrors (det < 0)."""
    out = transform_points(tr, pts)
    d = affine.det(tr)
    if out.ndim == 2:
        return out[FLIP_MAP] if d < 0 else out
    mask = d < 0
    out[mask] = out[mask][:, FLIP_MAP]
    return out


def transform_roi(tr, roi):
    """affinetrafo.py:75-86: map the 4 corners, take min / max."""
    roi = np.asarray(roi, F32)
    x0, y0, x1, y1 = np.moveaxis(roi, -1, 0)
    corners = np.stack(
        [np.stack([x0, y0], -1), np.stack([x0, y1], -1), np.stack([x1, y0], -1), np.stack([x1, y1], -1)], axis=-2
    )
    p = transform_points(tr, corners)
    return np.concatenate([p.min(axis=-2), p.max(axis=-2)], axis=-1).astype(F32)


def transform_coord(tr, coord):
    """affinetrafo.py:89-95: position through the map, size times the scale."""
    coord = np.asarray(coord, F32)
    out = np.empty_like(coord)
    out[..., :2] = affine.affinevecmul(tr, coord[..., :2])
    out[..., 2] = affine.scales(tr) * coord[..., 2]
    return out


def quat_mult(u, v):
    """torchquaternion.py:40-48, (i, j, k, w) order, Hamilton product u * v."""
    u, v = np.broadcast_arrays(np.asarray(u, F32), np.asarray(v, F32))
    ui, uj, uk, uw = np.moveaxis(u, -1, 0)
    vi, vj, vk, vw = np.moveaxis(v, -1, 0)
    w = uw * vw - ui * vi - uj * vj - uk * vk
    i = ui * vw + uw * vi - uk * vj + uj * vk
    j = uj * vw + uk * vi + uw * vj - ui * vk
    k = uk * vw - uj * vi + ui * vj + uw * vk
    return np.stack([i, j, k, w], axis=-1).astype(F32)


def quat_to_matrix(q):
    """torchquaternion.py:70-91 (the 6D-rotation target of losses.py:53-58 is its first two columns)."""
    q = np.asarray(q, F32)
    qi, qj, qk, qw = np.moveaxis(q, -1, 0)
    two, one = F32(2.0), F32(1.0)
    out = np.empty(q.shape[:-1] + (3, 3), F32)
    out[..., 0, 0] = one - two * (qj * qj + qk * qk)
    out[..., 1, 0] = two * (qi * qj + qk * qw)
    out[..., 2, 0] = two * (qi * qk - qj * qw)
    out[..., 0, 1] = two * (qi * qj - qk * qw)
    out[..., 1, 1] = one - two * (qi * qi + qk * qk)
    out[..., 2, 1] = two * (qj * qk + qi * qw)
    out[..., 0, 2] = two * (qi * qk + qj * qw)
    out[..., 1, 2] = two * (qj * qk - qi * qw)
    out[..., 2, 2] = one - two * (qi * qi + qj * qj)
    return out


def transform_rot(tr, quat):
    """affinetrafo.py:98-127: pre-multiply by the in-plane rotation of the map; mirror j, k when det < 0."""
    tr = np.asarray(tr, F32)
    quat = np.asarray(quat, F32)
    sn = -tr[..., 0, 1]
    cs = tr[..., 1, 1]
    detsign = np.sign(affine.det(tr)).astype(F32)
    alpha = np.arctan2(sn, cs).astype(F32)
    half = alpha * F32(0.5)
    qw = np.cos(half).astype(F32)
    qk = (np.sin(half).astype(F32) * detsign).astype(F32)
    zero = np.zeros_like(qw)
    zrot = np.stack([zero, zero, qk, qw], axis=-1)
    out = quat_mult(zrot, quat)
    out[..., 1] = detsign * out[..., 1]
    out[..., 2] = detsign * out[..., 2]
    return out


_TABLE = {CAT_XYS: transform_coord, CAT_QUAT: transform_rot, CAT_ROI: transform_roi, CAT_POINTS: transform_keypoints}


def apply_affine2d(tr, key, value, category):
    """affinetrafo.py:130-148: dispatch by field category; other fields pass through."""
    assert category not in IMAGELIKE
    if key == "image_backtransform":
        return affine.compose(np.asarray(value, F32), affine.inv(tr))
    fn = _TABLE.get(category)
    return value if fn is None else fn(tr, value)
