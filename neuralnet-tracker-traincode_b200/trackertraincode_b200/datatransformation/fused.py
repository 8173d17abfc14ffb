"""The whole training-time augmentation as ONE kernel launch per batch.

`FusedPoseAugmentation` is the drop-in for the reference's two transform chains taken together:
  per-sample chain  pipelines.py:372-383   offset_points_by_half_pixel -> [PutRoiFromLandmarks] -> RandomFocusRoi
                                           -> horizontal_flip_and_rot_90(0.01) -> normalize_batch
  per-batch chain   pipelines.py:508-532   to(device) -> KorniaImageDistortions x2 -> whiten_batch
It is meant to be installed as the loader's `postprocess` (datatransformation/loader.py:48-51) with workers that
return the *raw* sample (uint8 source frame + labels); frames of different sizes travel as a ragged list.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .. import _native as N
from ..datasets.batch import Batch
from . import _engine as E
from .batch.geometric import MakeRoiRandomizationParameters, NoRoiRandomization, draw_flip_rot90

# pipelines.py:510-527
OP_PROB = (0.2, 0.01, 0.2, 0.2, 0.2, 0.1)  # equalize, posterize, gamma, contrast, brightness, gaussian blur
RANDOM_APPLY = 4
NOISE_STD = (4.0 / 255.0, 16.0 / 255.0, 32.0 / 255.0, 64.0 / 255.0)
NOISE_PROB = (0.25, 0.25**2, 0.25**3, 0.25**4)


@dataclass
class AugmentationDraws:
    """All random numbers of one call (host tensors); pass to `FusedPoseAugmentation.__call__(params=...)` to replay."""

    geo: E.GeoParams
    do_flip: torch.Tensor
    rot_dir: torch.Tensor
    photo: Optional[E.PhotoParams]


def draw_photo_params(B: int, seed: int, sample_offset: int) -> E.PhotoParams:
    """kornia's per-call sampling for the two AugmentationSequential containers: which 4 of the 6 ops
    run (one multinomial draw per call), per-op per-sample Bernoulli masks, uniform factors."""
    order = torch.multinomial(torch.ones(N.NUM_OPS), RANDOM_APPLY).tolist()
    chosen = torch.zeros(N.NUM_OPS, dtype=torch.bool)
    chosen[order] = True
    apply = (torch.rand(B, N.NUM_OPS) < torch.tensor(OP_PROB)) & chosen
    u = torch.rand(4, B)
    return E.PhotoParams(
        order=order,
        apply=apply,
        bits=(4.0 + 2.0 * u[0]).to(torch.int32),
        gamma=0.5 + 1.5 * u[1],
        contrast=0.7 + 0.8 * u[2],
        brightness=0.7 + 0.8 * u[3],
        noise_apply=torch.rand(B, N.NUM_NOISE) < torch.tensor(NOISE_PROB),
        noise_std=NOISE_STD,
        seed=seed,
        sample_offset=sample_offset,
    )


class FusedPoseAugmentation:
    def __init__(self, inputsize: int = 129, rotation_aug_angle: float = 30.0, roi_override: str = "original",
                 enable_image_aug: bool = True, train: bool = True, p_rot: float = 0.01, device="cuda",
                 seed: int = 0, rowbuf_capacity: int = 0, zero_copy_frames: bool = False):
        if roi_override not in ("original", "landmarks"):
            raise N.NativeError("roi_override='extent_to_forehead' needs the BFM face model and is not on the B200 path")
        ext = {"original": 1.1, "landmarks": 1.2}[roi_override]  # pipelines.py:334
        self.inputsize = inputsize
        self.train = train
        self.p_rot = p_rot
        self.device = torch.device(device)
        self.enable_image_aug = enable_image_aug and train
        self.sampler = MakeRoiRandomizationParameters(rotation_aug_angle, ext) if train else NoRoiRandomization(ext)
        self.flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_NORMALIZE | N.F_WHITEN
        if roi_override == "landmarks":
            self.flags |= N.F_ROI_FROM_LANDMARKS
        if train:
            self.flags |= N.F_FLIPROT
        if self.enable_image_aug:
            self.flags |= N.F_PHOTOMETRIC
        self.seed = seed
        self.samples_seen = 0
        self.rowbuf_capacity = rowbuf_capacity
        self.zero_copy_frames = zero_copy_frames

    def draw(self, B: int) -> AugmentationDraws:
        p = self.sampler((B,))
        geo = E.GeoParams(p.scales, p.angles, p.translations, E.host_cos_sin(p.angles))
        if self.train:
            do_flip, rot_dir = draw_flip_rot90(self.p_rot, (B,))
        else:
            do_flip, rot_dir = torch.zeros(B, dtype=torch.uint8), torch.zeros(B, dtype=torch.int8)
        photo = draw_photo_params(B, self.seed, self.samples_seen) if self.enable_image_aug else None
        return AugmentationDraws(geo, do_flip, rot_dir, photo)

    def _to_device(self, batch: Batch) -> Batch:
        """Batch.to(device) (pipelines.py:508) -- except that source frames in PINNED host memory stay where they are: the
        kernel reads them in place over PCIe (zero copy), which moves only the bytes inside the view boxes (about 30 % of a
        450 x 450 frame at the pose pipeline's crop sizes) instead of whole frames.  Off by default: measured on B200 the
        SMs' small PCIe reads reach ~7 GB/s, the copy engine 50 GB/s, so copying whole frames is ~2x faster at config 2
        (profiles/README.md); it pays only when the view boxes are a small fraction of much larger frames."""
        if not self.zero_copy_frames:
            return batch.to(self.device, non_blocking=True)
        cats = batch.meta.categories

        def keep_on_host(k, v):
            if E.as_category(cats.get(k)) != E.FieldCategory.image:
                return False
            ts = v if isinstance(v, (list, tuple)) else [v]
            return all(isinstance(t, torch.Tensor) and t.dtype == torch.uint8 and not t.is_cuda and t.is_pinned() for t in ts)

        out = batch.__class__(batch.meta, {})
        for k, v in batch.items():
            if keep_on_host(k, v):
                out[k] = v
            elif isinstance(v, (list, tuple)):
                out[k] = [t.to(self.device, non_blocking=True) for t in v]
            else:
                out[k] = v.to(self.device, non_blocking=True)
        return out

    def __call__(self, batch: Batch, params: Optional[AugmentationDraws] = None) -> Batch:
        if batch.meta.prefixshape == ():
            batch = batch.with_batchdim()
        (B,) = batch.meta.prefixshape
        if batch.device != self.device:
            batch = self._to_device(batch)
        d = params if params is not None else self.draw(B)
        res = E.fused_forward(batch, flags=self.flags, out_size=self.inputsize, geo=d.geo, do_flip=d.do_flip,
                              rot_dir=d.rot_dir, photo=d.photo, rowbuf_capacity=self.rowbuf_capacity)
        self.samples_seen += B
        out = res.batch
        meta = batch.meta.__class__(**{**batch.meta.__dict__})
        meta._imagesize = self.inputsize
        out.meta = meta
        return out
