"""The whole training-time augmentation as ONE kernel launch per batch.

`FusedPoseAugmentation` is the drop-in for the reference's two transform chains taken together:
  per-sample chain  pipelines.py:372-383   offset_points_by_half_pixel -> [PutRoiFromLandmarks] -> RandomFocusRoi
                                           -> horizontal_flip_and_rot_90(0.01) -> normalize_batch
  per-batch chain   pipelines.py:508-532   to(device) -> KorniaImageDistortions x2 -> whiten_batch
It is meant to be installed as the loader's `postprocess` (datatransformation/loader.py:48-51) with workers that
return the *raw* sample (uint8 source frame + labels); frames of different sizes travel as a ragged list.
"""
from __future__ import annotations

import collections
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .. import _native as N
from ..datasets.batch import Batch
from . import _engine as E
from . import sharding
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
                 seed: int = 0, rowbuf_capacity: int = 0, zero_copy_frames: bool = False, upload_row_bands: bool = True,
                 headmodel=None):
        if roi_override not in ("original", "landmarks", "extent_to_forehead"):
            raise ValueError(f"got {roi_override}")  # pipelines.py:331
        ext = {"original": 1.1, "extent_to_forehead": 1.1, "landmarks": 1.2}[roi_override]  # pipelines.py:334
        # pipelines.py:352-356: the roi comes from the posed face model (needs the reference's BFM data), is not regenerated
        self.headmodel = None
        if roi_override == "extent_to_forehead":
            from ..facemodel import HeadModel

            self.headmodel = headmodel if headmodel is not None else HeadModel.default()
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
        self.upload_row_bands = upload_row_bands and not zero_copy_frames
        # whole row bands (1-D copies, the default) or only the boxes' columns (2-D copies: 34 % fewer bytes, but measured on
        # B200 the copy engine moves the ~270-byte rows of a batched 2-D copy at 8 GB/s against 45 GB/s for the bands --
        # 97 k vs 375 k samples/s end to end, profiles/README.md)
        self.upload_boxes = os.environ.get("B200AUG_UPLOAD_BOXES", "0") != "0"
        self.uploaded_rows = 0   # rows copied host->device by the last call that used the row-band upload
        self.uploaded_bytes = 0  # ... and the frame bytes
        self._frames = {}        # device frame stacks the row bands land in, per (shape, stream)
        self.steps = 0
        # Host buffers (and everything else the asynchronous launch points at) stay referenced until the stream has passed
        # the launch: the loader may drop its batch -- and its pin_memory thread reuse the pinned block -- while the copy
        # or the kernel is still queued.
        self._in_flight = collections.deque()
        self.status = E.StatusWatch()  # degenerate rois surface as NativeError one or two calls late (status.flush() waits)

    def _hold(self, stream, *objs):
        ev = torch.cuda.Event()
        ev.record(stream)
        self._in_flight.append((ev, objs))
        while self._in_flight and self._in_flight[0][0].query():
            self._in_flight.popleft()

    def draw(self, B: int) -> AugmentationDraws:
        p = self.sampler((B,))
        geo = E.GeoParams(p.scales, p.angles, p.translations, E.host_cos_sin(p.angles))
        if self.train:
            do_flip, rot_dir = draw_flip_rot90(self.p_rot, (B,))
        else:
            do_flip, rot_dir = torch.zeros(B, dtype=torch.uint8), torch.zeros(B, dtype=torch.int8)
        # the noise stream is keyed by a GLOBAL sample id: (step * world + rank) * local batch (sharding.py), so ranks never
        # share noise fields and a sample's noise does not depend on the world size
        rank, world = sharding.rank_world()
        offset = sharding.global_sample_offset(self.steps, rank, world, B)
        photo = draw_photo_params(B, self.seed, offset) if self.enable_image_aug else None
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

    @staticmethod
    def _account_for_video(meta, d: AugmentationDraws) -> AugmentationDraws:
        """geometric.py:180-191: every frame of a clip gets the crop draw of the clip's first frame; the flip / rot90 draw is
        per sample in the reference (geometric.py:234-237, applied before collation), i.e. also one per clip."""
        if getattr(meta, "seq", None) is None:
            return d
        g = d.geo
        for a, b in meta.sequence_start_end:
            g.translations[a:b, ...] = g.translations[a:a + 1, ...]
            g.scales[a:b] = g.scales[a:a + 1]
            if g.angles is not None:
                g.angles[a:b] = g.angles[a:a + 1]
            d.do_flip[a:b] = d.do_flip[a:a + 1]
            d.rot_dir[a:b] = d.rot_dir[a:a + 1]
        d.geo = E.GeoParams(g.scales, g.angles, g.translations, E.host_cos_sin(g.angles))
        return d

    # ---- row-band upload ------------------------------------------------------------------------------------
    def _touched_boxes(self, roi: torch.Tensor, d: AugmentationDraws, W: int, H: int, beyond_border_shift: float = 0.3) -> np.ndarray:
        """The part of each frame the kernel can touch, conservatively, as int32 [B,4] = x0, y0, x1, y1 clipped to the frame:
        the view box of geometric.py:135-156 restated in float64 with 3 pixels of margin (the kernel's float32 box differs by
        far less than a pixel); rotated samples read the bounding box of the rotated square, at most size / sqrt(2) either
        side of the box centre for any angle.  Columns are widened to multiples of 16 (the kernels fetch 16-byte vectors)."""
        r = roi.detach().to("cpu", torch.float64).reshape(-1, 4).numpy()
        f = d.geo.scales.detach().to("cpu", torch.float64).reshape(-1).numpy()
        t = d.geo.translations.detach().to("cpu", torch.float64).reshape(-1, 2).numpy()
        rot = d.geo.angles.detach().to("cpu", torch.float64).reshape(-1).numpy() != 0.0
        bw, bh = r[:, 2] - r[:, 0], r[:, 3] - r[:, 1]
        size = np.maximum(bw, bh) * f
        half = np.where(rot, size * 0.70711 + 1.0, size * 0.5)
        out = np.empty((r.shape[0], 4), np.int32)
        for axis, (extent, lim) in enumerate(((bw, W), (bh, H))):
            w = 0.5 * np.abs(size - extent) + beyond_border_shift * np.minimum(size, extent)
            c = 0.5 * (r[:, 2 + axis] + r[:, axis]) + w * t[:, axis]
            lo, hi = np.floor(c - half) - 3, np.ceil(c + half) + 3
            if axis == 0:
                lo, hi = np.floor(lo / 16.0) * 16.0, np.ceil(hi / 16.0) * 16.0
            lo = np.clip(lo, 0, lim)
            out[:, axis], out[:, 2 + axis] = lo, np.maximum(np.clip(hi, 0, lim), lo)
        return out

    def _upload(self, batch: Batch, d: AugmentationDraws) -> Batch:
        """Batch.to(device) (pipelines.py:508).  With `upload_row_bands` and stacked frames in pinned host memory only the
        rows the sampled view boxes touch are copied (b200aug_upload_row_bands): ~60 % of a 450 x 450 frame at the pose
        pipeline's crop sizes, and the host->device copy is what bounds the end-to-end rate."""
        cats = batch.meta.categories
        img_keys = [k for k, v in batch.items() if E.as_category(cats.get(k)) == E.FieldCategory.image]
        ok = (self.upload_row_bands and (self.flags & N.F_FOCUS) and not (self.flags & N.F_ROI_FROM_LANDMARKS) and len(img_keys) == 1
              and "roi" in batch.keys() and self.headmodel is None)  # (the bands follow the roi, which must be known on the host)
        img = batch[img_keys[0]] if ok else None
        ok = ok and isinstance(img, torch.Tensor) and img.dtype == torch.uint8 and not img.is_cuda and img.is_pinned() and img.is_contiguous() \
            and (img.dim() == 3 or (img.dim() == 4 and (img.shape[-1] == 1 or img.shape[1] == 1)))
        if not ok:
            return self._to_device(batch)
        B = img.shape[0]
        H, W = (img.shape[1], img.shape[2]) if (img.dim() == 3 or img.shape[-1] == 1) else (img.shape[2], img.shape[3])
        boxes = self._touched_boxes(batch["roi"], d, W, H)
        stream = torch.cuda.current_stream(self.device)
        key = (tuple(img.shape), stream.cuda_stream)
        frames = self._frames.get(key)
        if frames is None:
            frames = self._frames[key] = torch.empty(img.shape, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            if self.upload_boxes:
                N.check(N.lib.b200aug_upload_boxes(frames.data_ptr(), img.data_ptr(), H * W, W, B, boxes.ctypes.data, stream.cuda_stream),
                        "b200aug_upload_boxes")
                self.uploaded_bytes = int(((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).sum())
            else:
                lo, hi = np.ascontiguousarray(boxes[:, 1]), np.ascontiguousarray(boxes[:, 3])
                N.check(N.lib.b200aug_upload_row_bands(frames.data_ptr(), img.data_ptr(), H * W, W, B, lo.ctypes.data, hi.ctypes.data,
                                                       stream.cuda_stream), "b200aug_upload_row_bands")
                self.uploaded_bytes = int((hi - lo).sum()) * W
        self.uploaded_rows = int((boxes[:, 3] - boxes[:, 1]).sum())
        out = batch.__class__(batch.meta, {})
        for k, v in batch.items():
            out[k] = frames if k == img_keys[0] else (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else
                                                     [t.to(self.device, non_blocking=True) for t in v])
        return out

    def __call__(self, batch: Batch, params: Optional[AugmentationDraws] = None) -> Batch:
        if batch.meta.prefixshape == ():
            batch = batch.with_batchdim()
        (B,) = batch.meta.prefixshape
        d = params if params is not None else self._account_for_video(batch.meta, self.draw(B))
        src_batch = batch
        if batch.device != self.device:
            batch = self._upload(batch, d)
        if self.headmodel is not None and "pt3d_68" in batch:
            # PutRoiFromLandmarks(extend_to_forehead=True) behind offset_points_by_half_pixel (pipelines.py:352-353, 372): the
            # kernel applies the half-pixel shift to the labels itself, so the head box gets it as an argument
            batch = batch.__class__(batch.meta, dict(batch.items()))
            batch["roi"] = self.headmodel.roi(batch["coord"], batch["pose"], batch["shapeparam"] if "shapeparams" in batch else None,
                                              xy_offset=0.5)
        res = E.fused_forward(batch, flags=self.flags, out_size=self.inputsize, geo=d.geo, do_flip=d.do_flip,
                              rot_dir=d.rot_dir, photo=d.photo, rowbuf_capacity=self.rowbuf_capacity, want_status=True)
        self.status.watch(res.status, "FusedPoseAugmentation")
        self._hold(torch.cuda.current_stream(self.device), src_batch, batch, res)
        self.samples_seen += B
        self.steps += 1
        out = res.batch
        for k, v in src_batch.items():  # normalization.py:26-30: bool fields become smoothed 0.1 / 0.9 float32 targets
            if isinstance(v, torch.Tensor) and v.dtype == torch.bool:
                out[k] = torch.where(out[k].to(torch.bool), 0.9, 0.1).to(torch.float32)
        meta = batch.meta.__class__(**{**batch.meta.__dict__})
        meta._imagesize = self.inputsize
        out.meta = meta
        return out
