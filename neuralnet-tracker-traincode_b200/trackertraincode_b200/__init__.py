"""B200-native drop-in for the augmentation / label-transform path of opentrack/neuralnet-tracker-traincode.

Mirrors the reference's `trackertraincode.datatransformation` API (same callables, arguments and results) on
batched `Batch` objects living on a CUDA device; the arithmetic runs in hand-written sm_100a kernels behind the
C ABI of include/b200aug.h.  There is no CPU fallback.
"""
from . import datasets, neuralnets  # noqa: F401
from . import datatransformation  # noqa: F401

__all__ = ["datasets", "neuralnets", "datatransformation"]
