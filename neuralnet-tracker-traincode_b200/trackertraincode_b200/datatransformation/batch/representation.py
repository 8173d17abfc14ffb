"""to_numpy / to_tensor on whole batches, as trackertraincode/datatransformation/batch/representation.py:15-28.

Host-side container conversions (used by the dataset-creation notebooks and scripts/evaluate_pose_network.py:281-282), no
arithmetic.  `to_numpy` copies device tensors to the host first (the reference only ever sees CPU tensors here)."""
from __future__ import annotations

from copy import copy

import torch

from ...datasets.batch import Batch


def to_numpy(batch: Batch) -> Batch:
    """Convert batch from tensor to numpy array"""
    batch = copy(batch)
    for k, v in batch.items():
        batch[k] = [t.detach().cpu().numpy() for t in v] if isinstance(v, (list, tuple)) else v.detach().cpu().numpy()
    return batch


def to_tensor(batch: Batch) -> Batch:
    """Convert ndarrays in sample to Tensors."""
    batch = copy(batch)
    for k, v in batch.items():
        batch[k] = [torch.from_numpy(a) for a in v] if isinstance(v, (list, tuple)) else torch.from_numpy(v)
    return batch
