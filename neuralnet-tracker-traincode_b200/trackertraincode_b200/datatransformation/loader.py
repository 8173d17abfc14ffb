"""Loader glue, mirroring trackertraincode/datatransformation/loader.py:8-118 (same classes and arguments).

One extension for the B200 path: `ragged_images=True` makes the collation keep image fields as a python list of
per-sample uint8 tensors instead of stacking them, so DataLoader workers can return *raw* frames of different sizes and
the main process hands them to `FusedPoseAugmentation` (installed as `postprocess`), which crops them on the GPU.
"""
from __future__ import annotations

from typing import Any, Callable, Generator, Generic, TypeVar

from torch.utils.data import DataLoader, Dataset

from ..datasets.batch import Batch


class TransformedDataset(Dataset):
    def __init__(self, wrapped: Dataset, transform: Callable[[Batch], Batch]):
        super().__init__()
        self.transform = transform
        self.wrapped = wrapped

    def __len__(self):
        return len(self.wrapped)

    def __iter__(self) -> Generator[Batch, Any, None]:
        for x in self.wrapped:
            yield self.transform(x)

    def __getitem__(self, key) -> Batch:
        return self.transform(self.wrapped[key])


class SegmentedCollationDataLoader:
    def __init__(self, dataset: Dataset, *, batch_size: int, num_workers: int, segmentation_key_getter: Callable[[Batch], Any],
                 pin_memory: bool, sampler=None, worker_init_fn=None, postprocess: Callable[[Batch], Batch] = lambda x: x,
                 ragged_images: bool = False):
        self._loader = DataLoader(dataset=dataset, batch_size=batch_size, sampler=sampler, num_workers=num_workers,
                                  collate_fn=Batch.Collation(segmentation_key_getter, ragged_images=ragged_images),
                                  worker_init_fn=worker_init_fn, pin_memory=pin_memory)
        self._postprocess = postprocess

    def __iter__(self) -> Generator[list, Any, None]:
        for items in self._loader:
            assert isinstance(items, list)
            yield [self._postprocess(item) for item in items]

    def iter_unrolled(self) -> Generator[Batch, Any, None]:
        for items in self:
            yield from items

    def __len__(self):
        return len(self._loader)


T_co = TypeVar("T_co", covariant=True)


class PostprocessingLoader(Generic[T_co]):
    def __init__(self, *args, **kwargs):
        self._postprocess = kwargs.pop("postprocess", None) or (lambda x: x)
        self._loader = DataLoader(*args, **kwargs)

    @property
    def dataset(self) -> Dataset:
        return self._loader.dataset

    def __iter__(self):
        for items in self._loader:
            yield self._postprocess(items)

    def __len__(self):
        return len(self._loader)


class SampleBySampleLoader(Generic[T_co]):
    def __init__(self, dataset: Dataset, *, num_workers: int, pin_memory: bool = False, shuffle=False, sampler=None,
                 worker_init_fn=None, postprocess: Callable[[T_co], T_co] | None = None):
        self._loader = DataLoader(dataset=dataset, batch_size=max(1, num_workers), sampler=sampler, num_workers=num_workers,
                                  collate_fn=_identity_collate, worker_init_fn=worker_init_fn, pin_memory=pin_memory,
                                  shuffle=shuffle, drop_last=False)
        self._postprocess = postprocess or (lambda x: x)

    @property
    def dataset(self) -> Dataset:
        return self._loader.dataset

    def __iter__(self) -> Generator[T_co, Any, None]:
        for items in self._loader:
            assert isinstance(items, list)
            yield from (self._postprocess(item) for item in items)

    def __len__(self):
        return len(self._loader.dataset)


def _identity_collate(items):
    return items
