"""Device mirror of the quaternion helpers the label path uses (reference: neuralnets/torchquaternion.py).

Convention as in the reference: last dimension (i, j, k, w), real component last.  Tensors must live on a CUDA device;
the arithmetic runs in csrc/b200aug_fused.cu (quat_matrix_kernel) -- there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _native as N


def _run(x: torch.Tensor, in_tail, out_tail, to_matrix: int) -> torch.Tensor:
    if not x.is_cuda:
        raise N.NativeError(f"tensor lives on {x.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
    assert tuple(x.shape[-len(in_tail):]) == in_tail, f"expected trailing shape {in_tail}, got {tuple(x.shape)}"
    lead = x.shape[:-len(in_tail)]
    xin = x.to(torch.float32).reshape(-1, *in_tail).contiguous()
    out = torch.empty((xin.shape[0], *out_tail), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib.b200aug_quat_matrix(xin.data_ptr(), out.data_ptr(), xin.shape[0], to_matrix,
                                          torch.cuda.current_stream(x.device).cuda_stream), "b200aug_quat_matrix")
    return out.reshape(*lead, *out_tail)


def tomatrix(q: torch.Tensor) -> torch.Tensor:
    """torchquaternion.tomatrix (torchquaternion.py:70-91): normalised quaternions [..., 4] -> rotation matrices [..., 3, 3]
    (the target of losses.Rot6dReprLoss, neuralnets/losses.py:53-58)."""
    return _run(q, (4,), (3, 3), 1)


def from_matrix(m: torch.Tensor) -> torch.Tensor:
    """torchquaternion.from_matrix (torchquaternion.py:94-168): [..., 3, 3] -> [..., 4], real component made positive."""
    return _run(m, (3, 3), (4,), 0)
