"""The data model the transforms exchange: `Batch` (a dict of tensors with per-field categories) and `Metadata`.

API-compatible with trackertraincode/datasets/batch.py:15-238 (Metadata, Batch, Batch.Collation) and with
FieldCategory of trackertraincode/datasets/dshdf5pose.py:21-31, so code written against the reference (pipelines.py,
train_poseestimator.py, eval.py) keeps working.  One extension: an image-category field of a *collated* batch may be
a python list of per-sample tensors (ragged sources of different sizes) -- the reference itself does this in
eval.py:222-223 -- which lets un-cropped frames travel to the GPU where the fused kernel crops them.
"""
from __future__ import annotations

import copy as _copy
from collections import defaultdict
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple, Union

import numpy as np
import torch


class FieldCategory(str, Enum):
    general = ""
    image = "img"
    quat = "q"
    xys = "xys"
    roi = "roi"
    points = "pts"
    semseg = "seg"

    def __str__(self):
        return self.value


imagelike_categories = [FieldCategory.image, FieldCategory.semseg]


def as_category(c) -> FieldCategory:
    if isinstance(c, FieldCategory):
        return c
    if c is None:
        return FieldCategory.general
    return FieldCategory(str(c))


@dataclass
class Metadata:
    _imagesize: Union[int, Tuple[int, int]]
    batchsize: int
    tag: Optional[Any] = None
    seq: Optional[List[int]] = None
    categories: Dict[str, Any] = field(default_factory=dict)

    @property
    def image_wh(self) -> Tuple[int, int]:
        s = self._imagesize
        return s if isinstance(s, tuple) else (s, s)

    @property
    def imagesize(self) -> int:
        assert isinstance(self._imagesize, int)
        return self._imagesize

    @property
    def sequence_start_end(self):
        assert self.seq
        return list(zip(self.seq[:-1], self.seq[1:]))

    @property
    def prefixshape(self) -> tuple:
        if self.seq:
            return (self.seq[-1],)
        return (self.batchsize,) if self.batchsize else ()

    @property
    def is_single_frame(self) -> bool:
        return self.seq is None and self.batchsize == 0


def _cat(items):
    first = items[0]
    if isinstance(first, torch.Tensor):
        return torch.cat(list(items), dim=0)
    if isinstance(first, np.ndarray):
        return np.concatenate(list(items), axis=0)
    if isinstance(first, list):  # ragged image lists concatenate as lists
        flat = [x for it in items for x in it]
        if flat and all(isinstance(x, torch.Tensor) and x.dim() == 1 and x.dtype == torch.uint8 for x in flat):
            # encoded blobs (JPEG bytes of different lengths): ONE buffer, the list holds views of it -- a DataLoader worker
            # then ships one shared-memory segment per batch instead of one per sample
            packed = torch.cat(flat)
            ends = np.cumsum([int(x.numel()) for x in flat])
            return [packed[e - int(x.numel()):e] for x, e in zip(flat, ends)]
        return flat
    raise TypeError(f"cannot collate {type(first)}")


class Batch:
    def __init__(self, meta: Metadata, *data, **kwargs):
        self.meta = meta
        self._data: Dict[str, Any] = dict(*data, **kwargs)

    @staticmethod
    def from_data_with_categories(meta: Metadata, *args, **kwargs) -> "Batch":
        pairs = dict(*args, **kwargs)
        meta = _copy.copy(meta)
        meta.categories = dict(meta.categories)
        meta.categories.update((k, c) for k, (_, c) in pairs.items())
        return Batch(meta, ((k, v) for k, (v, _) in pairs.items()))

    # dict protocol ---------------------------------------------------------------------------------------
    def items(self):
        return self._data.items()

    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def __getitem__(self, k):
        return self._data[k]

    def __setitem__(self, k, v):
        self._data[k] = v

    def __delitem__(self, k):
        del self._data[k]

    def __contains__(self, k):
        return k in self._data

    def pop(self, k):
        return self._data.pop(k)

    def __copy__(self):
        return Batch(self.meta, self._data)

    def copy(self) -> "Batch":
        return Batch(self.meta, self._data)

    def __str__(self):
        return f"Batch({self.meta.tag},B={self.meta.batchsize})"

    @property
    def device(self):
        for v in self._data.values():
            if isinstance(v, torch.Tensor):
                return v.device
            if isinstance(v, list) and v and isinstance(v[0], torch.Tensor):
                return v[0].device
        raise ValueError("empty batch")

    def get_category(self, k, default=None):
        assert k in self._data
        return self.meta.categories.get(k, default)

    # views -----------------------------------------------------------------------------------------------
    def with_batchdim(self) -> "Batch":
        if self.meta.batchsize > 0:
            return self
        meta = _copy.copy(self.meta)
        meta.batchsize = 1
        if self.meta.seq is not None:
            return Batch(meta, self._data)
        return Batch(meta, ((k, [v] if _is_ragged_item(self, k, v) else v[None, ...]) for k, v in self._data.items()))

    def iter_frames(self) -> Iterator["Batch"]:
        if self.meta.is_single_frame:
            yield self
            return
        (n,) = self.meta.prefixshape
        meta = _copy.copy(self.meta)
        meta.batchsize, meta.seq = 0, None
        for i in range(n):
            yield Batch(meta, ((k, v[i]) for k, v in self._data.items()))

    def iter_sequences(self) -> Iterator["Batch"]:
        assert self.meta.seq is not None
        for a, b in self.meta.sequence_start_end:
            meta = _copy.copy(self.meta)
            meta.batchsize, meta.seq = 0, (0, b - a)
            yield Batch(meta, ((k, v[a:b]) for k, v in self._data.items()))

    def undo_collate(self):
        yield from (self.iter_sequences() if self.meta.seq else self.iter_frames())

    # movement ----------------------------------------------------------------------------------------------
    def _map_tensors(self, fn) -> "Batch":
        def one(v):
            if isinstance(v, torch.Tensor):
                return fn(v)
            if isinstance(v, list):
                return [fn(x) for x in v]
            raise AssertionError("Only applicable to PyTorch")

        return Batch(self.meta, ((k, one(v)) for k, v in self._data.items()))

    def pin_memory(self) -> "Batch":
        return self._map_tensors(lambda t: t.pin_memory())

    def to(self, *args, **kwargs) -> "Batch":
        return self._map_tensors(lambda t: t.to(*args, **kwargs))

    # collation ---------------------------------------------------------------------------------------------
    class Collation:
        """Concatenate samples into batches, optionally split by a key (the dataset tag in training)."""

        def __init__(self, key_getter: Optional[Callable[["Batch"], Any]] = None, ragged_images: bool = False):
            self._key_getter = key_getter if key_getter is not None else (lambda b: True)
            self._divide = key_getter is not None
            self._ragged = ragged_images

        def __call__(self, samples: List["Batch"]):
            groups = defaultdict(list)
            for s in samples:
                assert isinstance(s, Batch), f"Expected list of Batch types. Got {type(s)}"
                groups[self._key_getter(s)].append(s)
            out = [self._collate_group(g) for g in groups.values()]
            if not self._divide:
                (out,) = out
            return out

        def _collate_group(self, samples: List["Batch"]) -> "Batch":
            first = samples[0]
            meta = _copy.copy(first.meta)
            if first.meta.seq is None:
                meta.batchsize = sum(max(s.meta.batchsize, 1) for s in samples)
                parts = [self._with_batchdim(s) for s in samples]
            else:
                lengths = [s.meta.seq[-1] for s in samples]
                offsets = np.cumsum([0] + lengths[:-1])
                seq = [0]
                for s, o in zip(samples, offsets):
                    seq += [int(x + o) for x in s.meta.seq[1:]]
                meta.seq, meta.batchsize = seq, len(seq) - 1
                parts = samples
            data = {k: _cat([p[k] for p in parts]) for k in first.keys()}
            return Batch(meta, data)

        def _with_batchdim(self, s: "Batch") -> "Batch":
            if s.meta.batchsize > 0 or not self._ragged:
                return s.with_batchdim()
            meta = _copy.copy(s.meta)
            meta.batchsize = 1
            return Batch(meta, ((k, [v] if as_category(s.meta.categories.get(k)) in imagelike_categories else v[None, ...])
                                for k, v in s.items()))

    collate = Collation()


def _is_ragged_item(batch: Batch, k, v) -> bool:
    return isinstance(v, list)
