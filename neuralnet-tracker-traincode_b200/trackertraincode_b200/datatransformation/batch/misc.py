"""PutRoiFromLandmarks, mirroring trackertraincode/datatransformation/batch/misc.py:9-31."""
from __future__ import annotations

import copy as _copy

import torch

from ... import _native as N
from ...datasets.batch import Batch, FieldCategory, Metadata
from .. import _engine as E


class PutRoiFromLandmarks:
    """roi = [min_xy, max_xy] over the 68 landmarks (misc.py:22-25) or, with `extend_to_forehead=True`, over all vertices of
    the posed face model (misc.py:18-21).  The latter needs the reference's BFM data: `headmodel` (a facemodel.HeadModel), or
    HeadModel.default() which looks for the reference's pickle; it raises when there is none."""

    def __init__(self, extend_to_forehead: bool = False, headmodel=None):
        self.extend_to_forehead = bool(extend_to_forehead)
        self.headmodel = headmodel
        if self.extend_to_forehead and headmodel is None:
            from ...facemodel import HeadModel

            self.headmodel = HeadModel.default()

    def __call__(self, sample: Batch) -> Batch:
        if "pt3d_68" not in sample:
            return sample
        if self.extend_to_forehead:
            # misc.py:15-17 as written: the shape parameters are read only when the sample has a key "shapeparams" (the
            # datasets call it "shapeparam"), otherwise the mean shape is posed
            shape = sample["shapeparam"] if "shapeparams" in sample else None
            sample["roi"] = self.headmodel.roi(sample["coord"], sample["pose"], shape)
            sample.meta.categories = dict(sample.meta.categories, roi=sample.meta.categories.get("roi", FieldCategory.roi))
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
