"""ROI-focused random crop / scale / translate / rotate and flip / rot90, batched on the GPU.

Same callables and semantics as trackertraincode/datatransformation/batch/geometric.py:27-267; the difference is that
they accept whole batches (`meta.batchsize > 0`, per-sample parameters) as well as single frames, and that pixels and
labels are produced by the fused CUDA kernel instead of OpenCV + small torch ops.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np
import torch

from ... import _native as N
from ...datasets.batch import Batch, Metadata
from .. import _engine as E


class RoiFocusRandomizationParameters(NamedTuple):
    scales: torch.Tensor  # shape B
    angles: torch.Tensor  # shape B
    translations: torch.Tensor  # shape (B, 2)
    upfilter: Optional[str] = None
    downfilter: Optional[str] = None


class MakeRoiRandomizationParameters:
    """Sampler of geometric.py:58-84: scale ~ clip(0.1 N, +-0.5) + ext, translation ~ clip(0.5 N, +-1), angle = +-raug with
    probability 1/3.  Drawn on the host from the same torch / numpy global generators as the reference."""

    def __init__(self, rotation_aug_angle, extension_factor):
        self.rotation_aug_angle = rotation_aug_angle
        self.extension_factor = extension_factor

    def __call__(self, B: tuple) -> RoiFocusRandomizationParameters:
        scales = torch.randn(size=B).mul(0.1).clip(-0.5, 0.5).add(self.extension_factor)
        translations = torch.randn(size=B + (2,)).mul(0.5).clip(-1.0, 1.0)
        angles = self._pick_angles(B, self.rotation_aug_angle) if self.rotation_aug_angle else torch.zeros(size=B)
        return RoiFocusRandomizationParameters(scales, angles, translations, upfilter="linear", downfilter="area")

    @staticmethod
    def _pick_angles(B: tuple, angle: float):
        angles = torch.full(B, fill_value=np.pi * angle / 180.0)
        # single frames draw exactly like the reference (replace=False); a batch needs independent draws per sample
        rep = B != ()
        angles *= torch.from_numpy(np.asarray(np.random.choice([-1.0, 1.0], size=B, replace=rep)))
        angles *= torch.from_numpy(np.asarray(np.random.choice([0.0, 1.0], size=B, replace=rep, p=[2.0 / 3, 1.0 / 3])))
        return angles


class NoRoiRandomization:
    """geometric.py:87-96: the eval-time parameters."""

    def __init__(self, extent_factor):
        self.extent_factor = extent_factor

    def __call__(self, B) -> RoiFocusRandomizationParameters:
        return RoiFocusRandomizationParameters(torch.full(B, float(self.extent_factor)), torch.zeros(B), torch.zeros(B + (2,)))


def _check_filters(params: RoiFocusRandomizationParameters):
    """(downfilter, upfilter) for the kernels: 'area' | 'gaussian' | 'hamming' and 'linear' | 'cubic' | 'lanczos'
    (image_geometric_cv2.py:15-16, 47-82, 105-119); None = the defaults of :98-99."""
    up = params.upfilter or "linear"
    down = params.downfilter or "area"
    if up not in ("linear", "cubic", "lanczos"):
        raise KeyError(up)  # (the reference's dict lookup, image_geometric_cv2.py:70-75)
    if down not in ("area", "gaussian", "hamming"):
        raise NotImplementedError(f"Filter: {down}")
    return down, up


class GeneralFocusRoi:
    def __init__(self, make_randomization_parameters, new_size, roi_variable, insert_backtransform):
        self.new_size = new_size
        self.roi_variable = roi_variable
        self.insert_backtransform = insert_backtransform
        self._max_beyond_border_shift = 0.3
        self.make_randomization_parameters = make_randomization_parameters
        self.rowbuf_capacity = 0
        self.status = E.StatusWatch()  # empty view boxes surface as NativeError (deferred; status.flush() waits)

    @staticmethod
    def _maybe_account_for_video(meta: Metadata, params: RoiFocusRandomizationParameters):
        # geometric.py:180-191: every frame of a clip gets the draw of its first frame
        if meta.seq is None:
            return params
        for a, b in meta.sequence_start_end:
            params.translations[a:b, ...] = params.translations[a : a + 1, ...]
            params.scales[a:b] = params.scales[a : a + 1]
            if params.angles is not None:
                params.angles[a:b] = params.angles[a : a + 1]
        return params

    def __call__(self, sample: Batch) -> Batch:
        W, H = sample.meta.image_wh
        B = sample.meta.prefixshape
        params = self.make_randomization_parameters(B)
        downfilter, upfilter = _check_filters(params)
        self._maybe_account_for_video(sample.meta, params)
        geo = E.GeoParams(params.scales, params.angles, params.translations, E.host_cos_sin(params.angles))
        res = E.fused_forward(sample, flags=N.F_FOCUS, out_size=self.new_size, geo=geo, roi_variable=self.roi_variable,
                              beyond_border_shift=self._max_beyond_border_shift,
                              insert_backtransform=self.insert_backtransform, rowbuf_capacity=self.rowbuf_capacity,
                              want_status=True, downfilter=downfilter, upfilter=upfilter)
        self.status.watch(res.status, "GeneralFocusRoi")
        # like the reference, the passed sample (and its meta) is updated in place.  An "image_backtransform" that is
        # already there becomes BT @ tr^-1 (affinetrafo.py:137-147) -- unless insert_backtransform starts it afresh as tr^-1
        # (geometric.py:226-227); both come out of the kernel.
        for k, v in res.batch.items():
            sample[k] = v
        if self.insert_backtransform:
            size = torch.tensor((W, H), dtype=torch.int32, device=sample.device)
            sample["image_original_size"] = size.expand(*B, 2).contiguous() if B != () else size
        sample.meta._imagesize = self.new_size
        self.last_transform = res.tr
        return sample


def RandomFocusRoi(new_size, roi_variable="roi", rotation_aug_angle: float = 30.0, extension_factor=1.1, insert_backtransform=False):
    return GeneralFocusRoi(MakeRoiRandomizationParameters(rotation_aug_angle, extension_factor), new_size, roi_variable,
                           insert_backtransform)


def FocusRoi(new_size, extent_factor, roi_variable="roi", insert_backtransform=False):
    return GeneralFocusRoi(NoRoiRandomization(extent_factor), new_size, roi_variable, insert_backtransform)


def draw_flip_rot90(p_rot: float, B: tuple):
    """The two draws of geometric.py:236-237, per sample."""
    if B == ():
        do_flip = np.asarray(np.random.randint(0, 2) == 0)
        rot_dir = np.asarray(np.random.choice([-1, 0, 1], p=[p_rot / 2.0, (1.0 - p_rot), p_rot / 2.0]))
    else:
        do_flip = np.random.randint(0, 2, size=B) == 0
        rot_dir = np.random.choice([-1, 0, 1], size=B, p=[p_rot / 2.0, (1.0 - p_rot), p_rot / 2.0])
    return torch.from_numpy(do_flip.astype(np.uint8)), torch.from_numpy(rot_dir.astype(np.int8))


def horizontal_flip_and_rot_90(p_rot: float, sample: Batch, draws=None) -> Batch:
    """geometric.py:234-267.  `draws=(do_flip, rot_dir)` injects the random choices (tests / replay)."""
    B = sample.meta.prefixshape
    do_flip, rot_dir = draws if draws is not None else draw_flip_rot90(p_rot, B)
    do_flip, rot_dir = torch.as_tensor(do_flip), torch.as_tensor(rot_dir)
    if not bool(do_flip.any()) and not bool((rot_dir != 0).any()):
        return sample
    w, h = sample.meta.image_wh
    return E.fused_forward(sample, flags=N.F_FLIPROT, out_size=(w, h), do_flip=do_flip, rot_dir=rot_dir).batch
