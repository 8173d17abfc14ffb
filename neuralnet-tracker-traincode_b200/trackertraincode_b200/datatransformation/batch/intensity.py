"""Photometric augmentation containers, mirroring trackertraincode/datatransformation/batch/intensity.py:9-64.

The reference builds `KorniaImageDistortions(*kornia_ops, random_apply=N)` from kornia op objects (pipelines.py:510-527).
kornia is not a dependency here: the classes below carry the same constructor arguments as the kornia ops the reference
uses, only *describe* the op, and `KorniaImageDistortions.__call__` samples the per-call / per-sample parameters on the
host (the way kornia's AugmentationSequential does) and runs the whole container as ONE launch of
`b200aug_photometric_f32` (include/b200aug.h) per image field.  When the container follows the crop directly, prefer
`FusedPoseAugmentation`, which folds both stages into the crop kernel.
"""
from __future__ import annotations

import ctypes as C
from copy import copy
from typing import List, Optional, Sequence, Tuple

import torch

from ... import _native as N
from ...datasets.batch import Batch, FieldCategory, as_category
from .. import _engine as E
from .. import sharding


def _range(v, default_hi=None) -> Tuple[float, float]:
    if isinstance(v, (tuple, list)):
        return float(v[0]), float(v[1])
    return (float(v), float(v if default_hi is None else default_hi))


class _Op:
    op_id: Optional[int] = None

    def __init__(self, p: float = 0.5, same_on_batch: bool = False, keepdim: bool = False, **_ignored):
        if same_on_batch:
            raise N.NativeError("same_on_batch=True is not implemented on the B200 path")
        self.p = float(p)


class RandomEqualize(_Op):
    op_id = 0


class RandomPosterize(_Op):
    op_id = 1

    def __init__(self, bits=3, p: float = 0.5, **kw):
        super().__init__(p=p, **kw)
        self.bits = _range(bits) if isinstance(bits, (tuple, list)) else (float(bits), 8.0)


class RandomGamma(_Op):
    op_id = 2

    def __init__(self, gamma=(1.0, 1.0), gain=(1.0, 1.0), p: float = 1.0, **kw):
        super().__init__(p=p, **kw)
        self.gamma = _range(gamma)
        if _range(gain) != (1.0, 1.0):
            raise N.NativeError("RandomGamma gain != 1 is not implemented on the B200 path (pipelines.py uses the default)")


class RandomContrast(_Op):
    op_id = 3

    def __init__(self, contrast=(1.0, 1.0), clip_output: bool = True, p: float = 1.0, **kw):
        super().__init__(p=p, **kw)
        assert clip_output, "clip_output=False is not implemented"
        self.contrast = _range(contrast)


class RandomBrightness(_Op):
    op_id = 4

    def __init__(self, brightness=(1.0, 1.0), clip_output: bool = True, p: float = 1.0, **kw):
        super().__init__(p=p, **kw)
        assert clip_output, "clip_output=False is not implemented"
        self.brightness = _range(brightness)


class RandomGaussianBlur(_Op):
    op_id = 5

    def __init__(self, kernel_size=(5, 5), sigma=(1.5, 1.5), border_type: str = "reflect", separable: bool = True,
                 p: float = 0.5, silence_instantiation_warning: bool = False, **kw):
        super().__init__(p=p, **kw)
        ks = tuple(kernel_size) if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        if ks != (5, 5) or _range(sigma) != (1.5, 1.5) or border_type != "reflect":
            raise N.NativeError("the B200 blur is the reference's configuration only: 5x5, sigma 1.5, reflect (pipelines.py:516-518)")


class RandomGaussianNoise(_Op):
    def __init__(self, mean: float = 0.0, std: float = 1.0, p: float = 0.5, **kw):
        super().__init__(p=p, **kw)
        if mean != 0.0:
            raise N.NativeError("RandomGaussianNoise mean != 0 is not implemented on the B200 path")
        self.std = float(std)


class RandomGaussianNoiseWithClipping(RandomGaussianNoise):
    """intensity.py:43-53: the noise op whose output is clipped to [0, 1] (on the samples it was applied to)."""

    clip_output = True


class OnlyClip(_Op):
    """intensity.py:56-64: clip(0, 1) -- the reference applies it with p=1."""

    def __init__(self, p: float = 1.0, **kw):
        super().__init__(p=p, **kw)
        if self.p != 1.0:
            raise N.NativeError("OnlyClip with p != 1 is not implemented on the B200 path (pipelines.py:526 uses p=1.0)")


class KorniaImageDistortions:
    """Same call signature as the reference class; children must be the op descriptors of this module, stage-1 ops
    (each kind at most once) before the noise ops (at most 4) before an optional OnlyClip."""

    def __init__(self, *ops, random_apply: Optional[int] = None, seed: int = 0, bias: float = 0.0):
        self.point_ops: List[_Op] = []
        self.noise_ops: List[RandomGaussianNoise] = []
        self.clip = False
        for op in ops:
            if isinstance(op, OnlyClip):
                self.clip = True
            elif isinstance(op, RandomGaussianNoise):
                if self.clip:
                    raise N.NativeError("noise after OnlyClip is not a layout the B200 kernel implements")
                self.noise_ops.append(op)
            elif isinstance(op, _Op) and op.op_id is not None:
                if self.noise_ops or self.clip:
                    raise N.NativeError("point ops after noise/clip are not a layout the B200 kernel implements")
                if any(type(o) is type(op) for o in self.point_ops):
                    raise N.NativeError(f"{type(op).__name__} listed twice")
                self.point_ops.append(op)
            else:
                raise N.NativeError(f"unsupported child {op!r}: use the op descriptors of trackertraincode_b200.datatransformation.batch")
        if len(self.noise_ops) > N.NUM_NOISE:
            raise N.NativeError(f"at most {N.NUM_NOISE} noise stages")
        if random_apply is not None and (self.noise_ops or self.clip):
            raise N.NativeError("random_apply over noise / clip children is not implemented (the reference uses it on stage 1 only)")
        self.random_apply = random_apply
        self.seed = seed
        self.bias = bias
        self.samples_seen = 0
        self.calls = 0

    # -- sampling (host, torch global generator) ------------------------------------------------------------
    def draw(self, B: int) -> E.PhotoParams:
        n = len(self.point_ops)
        if self.random_apply is not None and n:
            pick = torch.multinomial(torch.ones(n), min(self.random_apply, n)).tolist()
        else:
            pick = list(range(n))
        order = [self.point_ops[i].op_id for i in pick]
        apply = torch.zeros(B, N.NUM_OPS, dtype=torch.bool)
        bits = torch.full((B,), 8, dtype=torch.int32)
        gamma, contrast, brightness = torch.ones(B), torch.ones(B), torch.ones(B)

        def uni(lo_hi):
            lo, hi = lo_hi
            return lo + (hi - lo) * torch.rand(B)

        for i in pick:
            op = self.point_ops[i]
            apply[:, op.op_id] = torch.rand(B) < op.p
            if isinstance(op, RandomPosterize):
                bits = uni(op.bits).to(torch.int32)  # kornia truncates the sampled float
            elif isinstance(op, RandomGamma):
                gamma = uni(op.gamma)
            elif isinstance(op, RandomContrast):
                contrast = uni(op.contrast)
            elif isinstance(op, RandomBrightness):
                brightness = uni(op.brightness)
        noise_apply = torch.zeros(B, N.NUM_NOISE, dtype=torch.bool)
        noise_std = [0.0] * N.NUM_NOISE
        noise_clip = [False] * N.NUM_NOISE
        for s, op in enumerate(self.noise_ops):
            noise_apply[:, s] = torch.rand(B) < op.p
            noise_std[s] = op.std
            noise_clip[s] = bool(getattr(op, "clip_output", False))
        # noise stream keyed by a global sample id (sharding.py): ranks never share noise fields
        rank, world = sharding.rank_world()
        offset = sharding.global_sample_offset(self.calls, rank, world, B)
        return E.PhotoParams(order, apply, bits, gamma, contrast, brightness, noise_apply, tuple(noise_std), self.seed,
                             offset, self.clip, tuple(noise_clip))

    def __call__(self, batch: Batch, params: Optional[E.PhotoParams] = None) -> Batch:
        batch = copy(batch)
        for k, v in batch.items():
            if as_category(batch.get_category(k)) != FieldCategory.image:
                continue
            batched = v.dim() == 4
            x = v if batched else v[None]
            p = params if params is not None else self.draw(x.shape[0])
            out = photometric_f32(x, p, bias=self.bias)
            batch[k] = out if batched else out[0]
            self.samples_seen += x.shape[0]
            self.calls += 1
        return batch


def photometric_f32(images: torch.Tensor, p: E.PhotoParams, bias: float = 0.0) -> torch.Tensor:
    """b200aug_photometric_f32 on float32 CUDA images [B, 1, H, W]; returns a new tensor."""
    if not images.is_cuda:
        raise N.NativeError(f"images live on {images.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
    if images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != 1:
        raise N.NativeError(f"expected float32 [B,1,H,W], got {images.dtype} {tuple(images.shape)}")
    x = images.contiguous()
    B, _, H, W = x.shape
    dev = x.device
    out = torch.empty_like(x)
    tmp = torch.empty_like(x) if N.OP_BLUR in list(p.order) else None
    pp, keep = E.marshal_photo(p, B, dev)
    with torch.cuda.device(dev):
        N.check(N.lib.b200aug_photometric_f32(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                             C.c_void_p(tmp.data_ptr() if tmp is not None else None), B, W, H, C.byref(pp),
                                             C.c_float(bias), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                "b200aug_photometric_f32")
    out._b200aug_keep = (keep, tmp, x)  # inputs must outlive the asynchronous launch
    return out
