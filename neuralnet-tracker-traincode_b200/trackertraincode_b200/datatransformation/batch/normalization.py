"""normalize / unnormalize / half-pixel offset / whiten on device batches.

Same functions as trackertraincode/datatransformation/batch/normalization.py:20-99.
"""
from __future__ import annotations

import copy as _copy

import torch

from ... import _native as N
from ...datasets.batch import Batch, FieldCategory, as_category, imagelike_categories
from .. import _engine as E


def _post_nonfloat(src: Batch, out: Batch) -> Batch:
    # normalization.py:26-30,45-48: bool -> smoothed 0.1/0.9 target, semseg -> long
    for k, v in src.items():
        if isinstance(v, torch.Tensor) and v.dtype == torch.bool:
            out[k] = torch.where(v, 0.9, 0.1).to(torch.float32)
    return out


def normalize_batch(sample: Batch) -> Batch:
    w, h = sample.meta.image_wh
    res = E.fused_forward(sample, flags=N.F_NORMALIZE, out_size=(w, h))
    return _post_nonfloat(sample, res.batch)


def unnormalize_batch(sample: Batch) -> Batch:
    """normalization.py:59-80: labels back to pixels, image clamp(x*256, 0, 255) -> uint8."""
    from ...neuralnets.affine2d import Affine2d

    w, h = sample.meta.image_wh
    tr = Affine2d.range_remap_2d([-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [w, h]).tensor().to(sample.device)
    out = _copy.copy(sample)
    names, fields = [], []
    batched = sample.meta.prefixshape != ()
    for k, v in sample.items():
        c = as_category(sample.meta.categories.get(k))
        if c == FieldCategory.image:
            out[k] = torch.clamp(v.mul(256.0), 0.0, 255.0).to(torch.uint8)
        elif k == "image_backtransform":  # BT @ tr^-1 (affinetrafo.py:137-147)
            names.append(k)
            fields.append(("image_backtransform", v if batched else v[None]))
        elif c in (FieldCategory.quat, FieldCategory.xys, FieldCategory.roi, FieldCategory.points):
            names.append(k)
            fields.append((c, v if batched else v[None]))
    if fields:
        for k, o in zip(names, E.apply_affine2d_fields(tr, fields)):
            out[k] = o if batched else o[0]
    return out


def offset_points_by_half_pixel(sample: Batch) -> Batch:
    """normalization.py:83-90: +0.5 on the xy of `pts` and `xys` fields (pixel centres)."""
    tr = torch.tensor([[1.0, 0.0, 0.5], [0.0, 1.0, 0.5]], device=sample.device)
    out = _copy.copy(sample)
    batched = sample.meta.prefixshape != ()
    names, fields = [], []
    for k, v in sample.items():
        c = as_category(sample.meta.categories.get(k))
        if c in (FieldCategory.points, FieldCategory.xys):
            names.append(k)
            fields.append((c, v if batched else v[None]))
    if fields:
        for k, o in zip(names, E.apply_affine2d_fields(tr, fields)):
            out[k] = o if batched else o[0]
    return out


def whiten_batch(batch: Batch) -> Batch:
    """normalization.py:94-99."""
    out = _copy.copy(batch)
    for k, v in batch.items():
        if as_category(batch.meta.categories.get(k)) in imagelike_categories:
            out[k] = v.sub(0.5)
    return out
