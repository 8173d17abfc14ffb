"""apply_affine2d and the position (un)normalisation transforms, tensors/affinetrafo.py:14-34,130-148.

`apply_affine2d(trafo, key, value, category)` transforms ONE label tensor with an explicit `Affine2d` through
b200aug_apply_affine2d (include/b200aug.h); the "image_backtransform" key is rewritten as BT @ trafo^-1 like the
reference does (:137-147).  eval.py:149-155 (`_apply_backtrafo`) is the consumer."""
from __future__ import annotations

import torch

from ... import _native as N
from ...datasets.batch import FieldCategory, as_category, imagelike_categories
from ...neuralnets.affine2d import Affine2d
from .. import _engine as E


def position_normalization(w: int, h: int) -> Affine2d:
    """affinetrafo.py:14-23: pixel coordinates -> [-1, 1]."""
    return Affine2d.range_remap_2d([0.0, 0.0], [w, h], [-1.0, -1.0], [1.0, 1.0])


def position_unnormalization(w: int, h: int) -> Affine2d:
    """affinetrafo.py:26-34: [-1, 1] -> pixel coordinates."""
    return Affine2d.range_remap_2d([-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [w, h])


_TRANSFORMED = (FieldCategory.xys, FieldCategory.quat, FieldCategory.roi, FieldCategory.points)


def apply_affine2d(trafo: Affine2d, key: str, value: torch.Tensor, category) -> torch.Tensor:
    category = as_category(category)
    assert category not in imagelike_categories
    is_bt = key == "image_backtransform"
    if not is_bt and category not in _TRANSFORMED:
        return value
    if not value.is_cuda:
        raise N.NativeError(f"{key!r} lives on {value.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
    tr = trafo.tensor().to(value.device, torch.float32)
    item_dims = 2 if is_bt else (2 if category == FieldCategory.points else 1)  # trailing dims of one sample's entry
    batched = tr.dim() == 3
    if batched:
        assert value.dim() >= item_dims + 1 and value.shape[0] == tr.shape[0], f"{key}: batch mismatch {tuple(value.shape)} vs {tuple(tr.shape)}"
        v = value
    else:
        # one transform for everything in `value`: flatten whatever prefix it has into the batch dimension
        lead = value.shape[: value.dim() - item_dims]
        v = value.reshape((-1,) + tuple(value.shape[value.dim() - item_dims:])) if len(lead) != 1 else value
    (out,) = E.apply_affine2d_fields(tr, [("image_backtransform" if is_bt else category, v)])
    return out.reshape(value.shape)
