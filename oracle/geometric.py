"""ROI-focused crop / rotate / scale of image + labels, and flip / rot90 -- TEST INFRASTRUCTURE.

Follows trackertraincode/datatransformation/batch/geometric.py:58-267 and
trackertraincode/datatransformation/tensors/image_geometric_cv2.py:28-155.  Pixels go through the same OpenCV
calls the reference makes (cv2 is the live ground truth); `use_model=True` routes them through the numpy
bit-models in oracle/cv2_model.py instead (identical output, used to prove the models).

A *sample* is a dict of numpy arrays plus a `categories` dict (field -> category string) and a `wh` tuple,
i.e. the reference's single-frame `Batch` (batchsize 0) without the class.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import cv2
import numpy as np

from . import affine, cv2_model
from .labels import CAT_IMAGE, IMAGELIKE, apply_affine2d

F32 = np.float32


@dataclass
class Sample:
    wh: Tuple[int, int]
    data: Dict[str, np.ndarray]
    categories: Dict[str, str] = field(default_factory=dict)

    def copy(self):
        return Sample(self.wh, dict(self.data), self.categories)


@dataclass
class RoiFocusParams:
    """geometric.py:27-32, one sample: scale (enlargement), angle [rad], translation factors in [-1, 1]^2."""

    scale: float
    angle: float
    translation: Tuple[float, float]
    cs_sn: Optional[Tuple[float, float]] = None  # optional (cos, sin) of the angle as the host computed them


def compute_view_roi(face_bbox, enlargement_factor, translation_factor, beyond_border_shift=0.3):
    """GeneralFocusRoi._compute_view_roi (geometric.py:107-157): float32 elementwise, un-rounded."""
    bb = np.asarray(face_bbox, F32)
    f = np.asarray(enlargement_factor, F32)
    t = np.asarray(translation_factor, F32)
    bbs = F32(beyond_border_shift)
    half = F32(0.5)
    x0, y0, x1, y1 = np.moveaxis(bb, -1, 0)
    rx, ry = np.moveaxis(t, -1, 0)
    bw = x1 - x0
    bh = y1 - y0
    cx = half * (x1 + x0)
    cy = half * (y1 + y0)
    size = np.maximum(bw, bh) * f
    wx = half * np.abs(size - bw) + bbs * np.minimum(size, bw)
    wy = half * np.abs(size - bh) + bbs * np.minimum(size, bh)
    tx = wx * rx
    ty = wy * ry
    hs = size * half
    return np.stack([cx - hs + tx, cy - hs + ty, cx + hs + tx, cy + hs + ty], axis=-1).astype(F32)


def round_view_roi(view_roi):
    """geometric.py:205: torch.round (half to even) -> int32."""
    return np.rint(np.asarray(view_roi, F32)).astype(np.int32)


def focus_transform(view_roi_i32, angle, new_wh, cs_sn=None):
    """geometric.py:159-178,206-207: tr = (denorm @ rot(angle) @ norm) @ range_remap(view_roi -> [0, new])."""
    ow, oh = new_wh
    vr = np.asarray(view_roi_i32).astype(F32)
    shape = vr.shape[:-1]
    tr_roi = affine.range_remap_2d(vr[..., :2], vr[..., 2:], np.zeros(shape + (2,), F32), np.broadcast_to(F32([ow, oh]), shape + (2,)))
    # NB the reference builds the centre rotation from (new_size, new_size); for non-square outputs we use (ow, oh).
    tr_norm = affine.position_normalization(ow, oh)
    tr_rot = affine.trs(angles=np.asarray(angle, F32), cs_sn=cs_sn)
    tr_denorm = affine.position_unnormalization(ow, oh)
    centre = affine.compose(affine.compose(tr_denorm, tr_rot), tr_norm)
    return affine.compose(centre, tr_roi)


# ----------------------------------------------------------------------------- pixels


def extract_roi_zero_padded(img: np.ndarray, roi) -> np.ndarray:
    """_numpy_extract_roi (image_geometric_cv2.py:28-44) for a [H, W] image: integer crop, zeros outside."""
    h, w = img.shape[:2]
    x0, y0, x1, y1 = (int(v) for v in roi)
    out = np.zeros((max(y1 - y0, 0), max(x1 - x0, 0)) + img.shape[2:], img.dtype)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0 : sy1 - y0, sx0 - x0 : sx1 - x0] = img[sy0:sy1, sx0:sx1]
    return out


def antialias_prefilter(img: np.ndarray, scale_factor: float, kind: str, use_model=False) -> np.ndarray:
    """_apply_antialias_filter (image_geometric_cv2.py:47-62): the cv2 calls, or their models."""
    if use_model:
        return cv2_model.antialias_prefilter_u8(img, scale_factor, kind)
    if kind == "gaussian":
        # the reference's call as Python binds it: sigmaX = ks, sigmaY = 1.0, default border (cv2_model.py explains)
        ks = 0.5 / scale_factor
        return cv2.GaussianBlur(img, (0, 0), sigmaX=ks, sigmaY=cv2_model.REFERENCE_GAUSSIAN_SIGMA_Y, borderType=cv2.BORDER_REFLECT_101)
    if kind == "hamming":
        kern = cv2_model.hamming_kernel(scale_factor)
        return cv2.sepFilter2D(img, -1, kern, kern)
    raise NotImplementedError(f"Filter: {kind}")


_CV2_UP = {"linear": cv2.INTER_LINEAR, "cubic": cv2.INTER_CUBIC, "lanczos": cv2.INTER_LANCZOS4}


def resize_area_or_linear(img: np.ndarray, new_w: int, new_h: int, use_model=False, downfilter="area", upfilter="linear") -> np.ndarray:
    """_resize (image_geometric_cv2.py:65-82): when the mean scale is < 1, INTER_AREA (downfilter='area', the sampler's fixed
    choice, geometric.py:76-77) or the gaussian / hamming prefilter followed by INTER_LINEAR; otherwise the up-filter
    (INTER_LINEAR -- the sampler's choice --, INTER_CUBIC or INTER_LANCZOS4)."""
    old_h, old_w = img.shape[:2]
    scale_factor = 0.5 * (new_w / old_w + new_h / old_h)
    if scale_factor >= 1.0 and upfilter != "linear":
        if use_model:
            return cv2_model.resize_cubic_or_lanczos_u8(img, new_w, new_h, upfilter)
        return cv2.resize(img, dsize=(new_w, new_h), interpolation=_CV2_UP[upfilter])
    if scale_factor < 1.0 and downfilter in ("gaussian", "hamming"):
        img = antialias_prefilter(img, scale_factor, downfilter, use_model)
        if use_model:
            return cv2_model.resize_linear_u8(img, new_w, new_h)
        return cv2.resize(img, dsize=(new_w, new_h), interpolation=cv2.INTER_LINEAR)
    area = scale_factor < 1.0
    if use_model:
        fn = cv2_model.resize_area_u8 if area else cv2_model.resize_linear_u8
        return fn(img, new_w, new_h)
    return cv2.resize(img, dsize=(new_w, new_h), interpolation=cv2.INTER_AREA if area else cv2.INTER_LINEAR)


def croprescale_image(img: np.ndarray, roi, new_wh, use_model=False, downfilter="area", upfilter="linear") -> np.ndarray:
    """croprescale_image_cv2 (image_geometric_cv2.py:138-155), [H, W] u8 in, [oh, ow] u8 out."""
    ow, oh = new_wh
    return resize_area_or_linear(extract_roi_zero_padded(img, roi), ow, oh, use_model, downfilter, upfilter)


def warp_plan(tr, new_wh):
    """The decisions affine_transform_image_cv2 (image_geometric_cv2.py:85-135) takes before touching pixels.

    Returns (M float32 2x3 handed to cv2.warpAffine, canvas_w, canvas_h, upscale: bool).
    """
    ow, oh = new_wh
    tr = np.asarray(tr, F32)
    scale_factor = float(affine.scales(tr))
    if scale_factor > 1.0:
        M = affine.compose(tr, affine.trs(translations=F32([0.5, 0.5])))
        return M, ow, oh, True
    rot_w, rot_h = round(ow / scale_factor), round(oh / scale_factor)  # Python banker's rounding on doubles
    scale_compensation = F32(rot_h / oh)
    M = affine.compose(affine.trs(scales=scale_compensation), tr)
    return M, rot_w, rot_h, False


def affine_transform_image(img: np.ndarray, tr, new_wh, use_model=False, downfilter="area", upfilter="linear") -> np.ndarray:
    """affine_transform_image_cv2: an up-scaling transform warps straight to the output with the up-filter (:105-119);
    otherwise anti-aliased = bilinear warp to an intermediate canvas at source resolution, then _resize (:120-134)."""
    ow, oh = new_wh
    M, cw, ch, up = warp_plan(tr, new_wh)
    interp = upfilter if up else "linear"
    if use_model:
        canvas = (cv2_model.warp_affine_linear_u8(img, M, cw, ch) if interp == "linear"
                  else cv2_model.warp_affine_cubic_or_lanczos_u8(img, M, cw, ch, interp))
    else:
        canvas = cv2.warpAffine(img, M=M, dsize=(cw, ch), flags=_CV2_UP[interp], borderMode=cv2.BORDER_CONSTANT, borderValue=None)
    if up:
        return canvas
    return resize_area_or_linear(canvas, ow, oh, use_model, downfilter, upfilter)


# ----------------------------------------------------------------------------- sample-level transforms


def focus_roi(sample: Sample, params: RoiFocusParams, new_size, roi_variable="roi", insert_backtransform=False,
              beyond_border_shift=0.3, use_model=False, downfilter="area", upfilter="linear") -> Tuple[Sample, dict]:
    """GeneralFocusRoi.__call__ (geometric.py:193-231) with explicit parameters.

    Returns the transformed sample and the intermediates the parity tests compare bit-exactly
    ({'view_roi': int32[4], 'tr': float32[2,3]}).  Image out: uint8 [1, oh, ow] (CHW like the reference)."""
    new_wh = (new_size, new_size) if isinstance(new_size, int) else tuple(new_size)
    W, H = sample.wh
    roi = sample.data[roi_variable]
    view = compute_view_roi(roi, F32(params.scale), F32(params.translation), beyond_border_shift)
    view_i = round_view_roi(view)
    tr = focus_transform(view_i, F32(params.angle), new_wh, params.cs_sn)
    out = sample.copy()
    for k, v in sample.data.items():
        c = sample.categories.get(k, "")
        if c == CAT_IMAGE:
            img = v[..., 0] if v.ndim == 3 else v
            if F32(params.angle) != 0.0:
                res = affine_transform_image(img, tr, new_wh, use_model, downfilter, upfilter)
            else:
                res = croprescale_image(img, view_i, new_wh, use_model, downfilter, upfilter)
            out.data[k] = res[None, ...]
        elif c in IMAGELIKE:
            raise NotImplementedError("semseg fields are outside the hot path")
        else:
            out.data[k] = apply_affine2d(tr, k, v, c)
    if insert_backtransform:
        out.data["image_backtransform"] = affine.inv(tr)
        out.data["image_original_size"] = np.asarray((W, H), np.int32)
    out.wh = new_wh
    return out, {"view_roi": view_i, "tr": tr}


def flip_rot90_transform(do_flip: bool, rot_dir: int, wh):
    """The label transform of horizontal_flip_and_rot_90 (geometric.py:242-252)."""
    w, h = wh
    tr = affine.identity()
    if rot_dir != 0:
        tr = affine.compose(tr, affine.range_remap_2d([-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [w, h]))
        tr = affine.compose(tr, affine.trs(angles=F32(rot_dir * np.pi * 0.5)))
        tr = affine.compose(tr, affine.range_remap_2d([0.0, 0.0], [w, h], [-1.0, -1.0], [1.0, 1.0]))
    if do_flip:
        tr = affine.compose(tr, affine.range_remap_2d([0.0, 0.0], [w, h], [w, 0], [0, h]))
    return tr


def flip_rot90_image(v: np.ndarray, do_flip: bool, rot_dir: int) -> np.ndarray:
    """geometric.py:256-264 on a [..., H, W] array: pure permutation."""
    if do_flip:
        v = v[..., ::-1]
    if rot_dir != 0:
        v = np.swapaxes(v, -1, -2)
    if rot_dir == 1:
        v = v[..., ::-1]
    elif rot_dir == -1:
        v = v[..., ::-1, :]
    return np.ascontiguousarray(v)


def horizontal_flip_and_rot_90(sample: Sample, do_flip: bool, rot_dir: int) -> Sample:
    """horizontal_flip_and_rot_90 (geometric.py:234-267) with the two random draws made explicit."""
    if not do_flip and rot_dir == 0:
        return sample
    out = sample.copy()
    tr = flip_rot90_transform(do_flip, rot_dir, sample.wh)
    for k, v in sample.data.items():
        c = sample.categories.get(k, "")
        if c in IMAGELIKE:
            out.data[k] = flip_rot90_image(v, do_flip, rot_dir)
        else:
            out.data[k] = apply_affine2d(tr, k, v, c)
    return out


def put_roi_from_landmarks(sample: Sample) -> Sample:
    """PutRoiFromLandmarks(extend_to_forehead=False) (batch/misc.py:9-31): roi = [min_xy, max_xy] of pt3d_68."""
    if "pt3d_68" not in sample.data:
        return sample
    out = sample.copy()
    p = np.asarray(sample.data["pt3d_68"], F32)
    out.data["roi"] = np.concatenate([p[..., :2].min(axis=-2), p[..., :2].max(axis=-2)], axis=-1).astype(F32)
    out.categories = dict(sample.categories, roi="roi")
    return out
