"""normalize / unnormalize / half-pixel offset / whiten -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows trackertraincode/datatransformation/batch/normalization.py:20-99 and
trackertraincode/datatransformation/tensors/normalization.py:19-24.
"""
from __future__ import annotations

import numpy as np

from . import affine
from .geometric import Sample
from .labels import CAT_IMAGE, CAT_POINTS, CAT_SEMSEG, CAT_XYS, IMAGELIKE, apply_affine2d

F32 = np.float32


def normalize_image(x: np.ndarray) -> np.ndarray:
    """normalization.py:32-34: uint8 -> float32 * (1/256)  (256, not 255)."""
    return x.astype(F32) * F32(1.0 / 256)


def normalize_bool(x: np.ndarray, smooth=0.1) -> np.ndarray:
    """normalization.py:26-30: label smoothing, False -> 0.1, True -> 0.9."""
    out = np.full(x.shape, F32(smooth), F32)
    out[x] = F32(1.0 - smooth)
    return out


def normalize_sample(sample: Sample) -> Sample:
    """normalize_batch (normalization.py:20-56)."""
    w, h = sample.wh
    tr = affine.position_normalization(w, h)
    out = sample.copy()
    for k, v in sample.data.items():
        c = sample.categories.get(k, "")
        if c == CAT_IMAGE:
            out.data[k] = normalize_image(v)
        elif c == CAT_SEMSEG:
            out.data[k] = v.astype(np.int64)
        elif v.dtype == np.bool_:
            out.data[k] = normalize_bool(v)
        else:
            out.data[k] = apply_affine2d(tr, k, v, c)
    return out


def unnormalize_image(x: np.ndarray) -> np.ndarray:
    """normalization.py:66-68: clamp(x*256, 0, 255) -> uint8 (truncation)."""
    return np.clip(x.astype(F32) * F32(256.0), 0.0, 255.0).astype(np.uint8)


def unnormalize_sample(sample: Sample) -> Sample:
    """unnormalize_batch (normalization.py:59-80)."""
    w, h = sample.wh
    tr = affine.position_unnormalization(w, h)
    out = sample.copy()
    for k, v in sample.data.items():
        c = sample.categories.get(k, "")
        if c == CAT_IMAGE:
            out.data[k] = unnormalize_image(v)
        else:
            out.data[k] = apply_affine2d(tr, k, v, c)
    return out


def offset_points_by_half_pixel(sample: Sample) -> Sample:
    """normalization.py:83-90: +0.5 px on landmark and head-coordinate positions (pixel centres)."""
    tr = affine.trs(translations=F32([0.5, 0.5]))
    out = sample.copy()
    for k, v in sample.data.items():
        c = sample.categories.get(k, "")
        if c in (CAT_POINTS, CAT_XYS):
            out.data[k] = apply_affine2d(tr, k, v, c)
    return out


def whiten_image(x: np.ndarray) -> np.ndarray:
    """tensors/normalization.py:19-20."""
    return x - F32(0.5)


def whiten_sample(sample: Sample) -> Sample:
    """whiten_batch (normalization.py:94-99)."""
    out = sample.copy()
    for k, v in sample.data.items():
        if sample.categories.get(k, "") in IMAGELIKE:
            out.data[k] = whiten_image(v)
    return out
