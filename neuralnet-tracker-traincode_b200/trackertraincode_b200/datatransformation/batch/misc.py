"""PutRoiFromLandmarks, mirroring trackertraincode/datatransformation/batch/misc.py:9-31."""
from __future__ import annotations

import copy as _copy

import torch

from ... import _native as N
from ...datasets.batch import Batch, FieldCategory, Metadata
from .. import _engine as E


class PutRoiFromLandmarks:
    """roi = [min_xy, max_xy] over the 68 landmarks (misc.py:22-25).  `extend_to_forehead=True` needs the BFM head model
    (misc.py:12,18-21), which is outside this path: it raises instead of silently doing something else."""

    def __init__(self, extend_to_forehead: bool = False):
        if extend_to_forehead:
            raise N.NativeError("extend_to_forehead=True needs the BFM face model and is not on the B200 path")
        self.extend_to_forehead = False

    def __call__(self, sample: Batch) -> Batch:
        if "pt3d_68" not in sample:
            return sample
        pts = sample["pt3d_68"]
        batched = sample.meta.prefixshape != ()
        B = sample.meta.prefixshape[0] if batched else 1
        roi_in = torch.zeros((B, 4) if batched else (4,), dtype=torch.float32, device=pts.device)
        sub = Batch(Metadata(sample.meta._imagesize, sample.meta.batchsize, sample.meta.tag, sample.meta.seq,
                             {"pt3d_68": FieldCategory.points, "roi": FieldCategory.roi}), {"pt3d_68": pts, "roi": roi_in})
        res = E.fused_forward(sub, flags=N.F_ROI_FROM_LANDMARKS, out_size=sample.meta.image_wh)
        sample["roi"] = res.batch["roi"]
        sample.meta.categories = dict(sample.meta.categories, roi=sample.meta.categories.get("roi", FieldCategory.roi))
        return sample
