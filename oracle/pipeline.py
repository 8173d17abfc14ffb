"""The training-time augmentation chain end to end -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Per-sample half (runs in DataLoader workers in the reference), trackertraincode/pipelines.py:372-383:
    offset_points_by_half_pixel -> [PutRoiFromLandmarks] -> RandomFocusRoi -> [PutRoiFromLandmarks]
    -> horizontal_flip_and_rot_90 -> normalize_batch
Per-batch half (the loader's `postprocess`), trackertraincode/pipelines.py:508-532:
    KorniaImageDistortions x2 -> whiten_batch
All random draws are explicit inputs so the CUDA path can be fed identical parameters.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import geometric as geo
from . import normalization as nrm
from . import photometric as pho
from .geometric import RoiFocusParams, Sample

F32 = np.float32


@dataclass
class GeoParams:
    """Batch of geometric draws: RoiFocusRandomizationParameters (batch/geometric.py:27-32) + flip/rot90 draws."""

    scales: np.ndarray  # float32 [B]
    angles: np.ndarray  # float32 [B], radians
    translations: np.ndarray  # float32 [B, 2]
    do_flip: np.ndarray  # bool [B]
    rot_dir: np.ndarray  # int8 [B] in {-1, 0, 1}

    def __len__(self):
        return len(self.scales)

    def slice(self, lo, hi):
        return GeoParams(self.scales[lo:hi], self.angles[lo:hi], self.translations[lo:hi], self.do_flip[lo:hi], self.rot_dir[lo:hi])


def sample_geo_params(rng: np.random.Generator, B: int, rotation_aug_angle=30.0, extension_factor=1.1, p_rot=0.01) -> GeoParams:
    """MakeRoiRandomizationParameters.__call__/_pick_angles (batch/geometric.py:63-84) and the two draws of
    horizontal_flip_and_rot_90 (batch/geometric.py:236-237), as distributions (not the torch/numpy RNG streams)."""
    scales = (np.clip(F32(0.1) * rng.standard_normal(B).astype(F32), -0.5, 0.5) + F32(extension_factor)).astype(F32)
    translations = np.clip(F32(0.5) * rng.standard_normal((B, 2)).astype(F32), -1.0, 1.0).astype(F32)
    if rotation_aug_angle:
        angles = np.full(B, np.pi * rotation_aug_angle / 180.0, F32)
        angles *= rng.choice(np.asarray([-1.0, 1.0], F32), size=B)
        angles *= rng.choice(np.asarray([0.0, 1.0], F32), size=B, p=[2.0 / 3, 1.0 / 3])
    else:
        angles = np.zeros(B, F32)
    do_flip = rng.integers(0, 2, B) == 0
    rot_dir = rng.choice(np.asarray([-1, 0, 1], np.int8), size=B, p=[p_rot / 2.0, 1.0 - p_rot, p_rot / 2.0])
    return GeoParams(scales, angles, translations, do_flip, rot_dir.astype(np.int8))


def no_randomization(B: int, extent_factor: float) -> GeoParams:
    """NoRoiRandomization (batch/geometric.py:87-96) + no flip: the eval-time transform."""
    return GeoParams(np.full(B, extent_factor, F32), np.zeros(B, F32), np.zeros((B, 2), F32), np.zeros(B, bool), np.zeros(B, np.int8))


def augment_sample(sample: Sample, scale, angle, translation, do_flip, rot_dir, new_size=129, roi_mode="original",
                   use_model=False, half_pixel=True, normalize=True, downfilter="area", upfilter="linear"):
    """The per-sample half for one sample; returns (Sample, intermediates)."""
    s = nrm.offset_points_by_half_pixel(sample) if half_pixel else sample
    if roi_mode == "landmarks":
        s = geo.put_roi_from_landmarks(s)
    elif roi_mode != "original":
        raise NotImplementedError("extent_to_forehead needs the BFM face model; out of the hot-path scope")
    s, inter = geo.focus_roi(s, RoiFocusParams(scale, angle, tuple(translation)), new_size, use_model=use_model, downfilter=downfilter,
                              upfilter=upfilter)
    if roi_mode == "landmarks":
        s = geo.put_roi_from_landmarks(s)
    s = geo.horizontal_flip_and_rot_90(s, bool(do_flip), int(rot_dir))
    if normalize:
        s = nrm.normalize_sample(s)
    return s, inter


def collate(samples: Sequence[Sample]) -> Dict[str, np.ndarray]:
    """Batch.Collation for stills (trackertraincode/datasets/batch.py:199-236): stack every field."""
    return {k: np.stack([np.asarray(s.data[k]) for s in samples], axis=0) for k in samples[0].data}


def augment_batch(samples: Sequence[Sample], gp: GeoParams, pp: Optional[pho.PhotoParams], new_size=129,
                  roi_mode="original", use_model=False, whiten=True, downfilter="area", upfilter="linear"):
    """Full chain on a list of source samples.  Returns (dict of stacked outputs, dict of stacked intermediates)."""
    outs, inters = [], []
    for i, s in enumerate(samples):
        o, it = augment_sample(s, gp.scales[i], gp.angles[i], gp.translations[i], gp.do_flip[i], gp.rot_dir[i],
                               new_size, roi_mode, use_model, downfilter=downfilter, upfilter=upfilter)
        outs.append(o)
        inters.append(it)
    batch = collate(outs)
    cats = outs[0].categories
    for k, c in cats.items():
        if c == "img" and k in batch:
            x = batch[k]
            if pp is not None:
                x = pho.photometric_batch(x, pp)
            if whiten:
                x = nrm.whiten_image(x)
            batch[k] = x
    inter = {k: np.stack([it[k] for it in inters], 0) for k in inters[0]}
    return batch, inter
