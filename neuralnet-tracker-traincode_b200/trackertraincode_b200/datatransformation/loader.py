"""Loader glue with the names and call signatures of trackertraincode/datatransformation/loader.py:8-118.

The three loaders of the reference differ only in how worker output is grouped (segmented `Batch` lists, whatever the
caller's `collate_fn` makes, single samples) and all end in a `postprocess` callable run in the main process -- the
place where `FusedPoseAugmentation` is installed (pipelines.py:508-543).  Here they share one base, `_Piped`, which owns
the torch `DataLoader` and the postprocess hook; each public class only says how to build the `DataLoader` and how to
unpack what it yields.

Extensions for the B200 path (both default off, i.e. reference behaviour):
  * `ragged_images=True` (`SegmentedCollationDataLoader`): the collation keeps image fields as per-sample uint8 tensors
    (one packed buffer per batch) instead of stacking them, so workers can return raw frames or JPEG blobs of different
    sizes and the crop happens on the GPU;
  * `lookahead=n`: the postprocess of the next `n` groups is issued before the current group is handed out, so its
    host-side marshalling and H2D copies overlap the consumer's GPU work (the postprocess must be asynchronous with
    respect to the device, as `FusedPoseAugmentation` is).
"""
from __future__ import annotations

from collections import deque
from typing import Any, Callable, Generic, Iterator, TypeVar

from torch.utils.data import DataLoader, Dataset

from ..datasets.batch import Batch

T_co = TypeVar("T_co", covariant=True)


def _same(x):
    return x


def _as_list(items):
    return items


class TransformedDataset(Dataset):
    """`wrapped` seen through `transform` (loader.py:8-24): indexable and iterable like the dataset underneath."""

    def __init__(self, wrapped: Dataset, transform: Callable[[Batch], Batch]):
        super().__init__()
        self.wrapped, self.transform = wrapped, transform

    def __getitem__(self, key) -> Batch:
        return self.transform(self.wrapped[key])

    def __iter__(self) -> Iterator[Batch]:
        return map(self.transform, iter(self.wrapped))

    def __len__(self):
        return len(self.wrapped)


class _Piped:
    """A torch DataLoader followed by a main-process hook.  Subclasses implement `_emit(group)`: an iterator over the
    postprocessed things one DataLoader item turns into."""

    def __init__(self, inner: DataLoader, postprocess, lookahead: int = 0):
        self._loader = inner
        self._postprocess = postprocess if postprocess is not None else _same
        self._lookahead = max(0, int(lookahead))

    @property
    def dataset(self) -> Dataset:
        return self._loader.dataset

    def _emit(self, group):
        raise NotImplementedError

    def _groups(self):
        """Postprocessed DataLoader items in order, `lookahead` of them issued ahead of the one being consumed."""
        pending: deque = deque()
        for group in self._loader:
            pending.append(self._emit(group))
            if len(pending) > self._lookahead:
                yield pending.popleft()
        while pending:
            yield pending.popleft()

    def __len__(self):
        return len(self._loader)


class SegmentedCollationDataLoader(_Piped):
    """Samples of a batch are grouped by `segmentation_key_getter` and every group is collated to its own `Batch`; one
    iteration step yields the list of postprocessed groups (loader.py:27-58)."""

    def __init__(self, dataset: Dataset, *, batch_size: int, num_workers: int, segmentation_key_getter: Callable[[Batch], Any],
                 pin_memory: bool, sampler=None, worker_init_fn=None, postprocess: Callable[[Batch], Batch] | None = None,
                 ragged_images: bool = False, lookahead: int = 0):
        collate = Batch.Collation(segmentation_key_getter, ragged_images=ragged_images)
        super().__init__(DataLoader(dataset=dataset, batch_size=batch_size, sampler=sampler, num_workers=num_workers, collate_fn=collate,
                                    worker_init_fn=worker_init_fn, pin_memory=pin_memory), postprocess, lookahead)

    def _emit(self, group) -> list:
        if not isinstance(group, list):
            raise TypeError(f"the segmented collation yields lists of Batch, got {type(group).__name__}")
        return [self._postprocess(segment) for segment in group]

    def __iter__(self) -> Iterator[list]:
        return self._groups()

    def iter_unrolled(self) -> Iterator[Batch]:
        for segments in self._groups():
            yield from segments


class PostprocessingLoader(_Piped, Generic[T_co]):
    """A plain `DataLoader(*args, **kwargs)` whose items go through `postprocess` (loader.py:64-80)."""

    def __init__(self, *args, postprocess=None, lookahead: int = 0, **kwargs):
        super().__init__(DataLoader(*args, **kwargs), postprocess, lookahead)

    def _emit(self, group):
        return self._postprocess(group)

    def __iter__(self):
        return self._groups()


class SampleBySampleLoader(_Piped, Generic[T_co]):
    """Uncollated samples one at a time; workers hand them over `num_workers` at a go (loader.py:83-118).  `len()` is the
    number of samples, not of DataLoader steps."""

    def __init__(self, dataset: Dataset, *, num_workers: int, pin_memory: bool = False, shuffle=False, sampler=None,
                 worker_init_fn=None, postprocess: Callable[[T_co], T_co] | None = None, lookahead: int = 0):
        super().__init__(DataLoader(dataset=dataset, batch_size=max(1, num_workers), num_workers=num_workers, shuffle=shuffle, sampler=sampler,
                                    drop_last=False, collate_fn=_as_list, worker_init_fn=worker_init_fn, pin_memory=pin_memory),
                         postprocess, lookahead)

    def _emit(self, group) -> list:
        if not isinstance(group, list):
            raise TypeError(f"expected a list of samples, got {type(group).__name__}")
        return [self._postprocess(sample) for sample in group]

    def __iter__(self) -> Iterator[T_co]:
        for samples in self._groups():
            yield from samples

    def __len__(self):
        return len(self._loader.dataset)
