"""Device mirror of the rotation post-processing the evaluation side applies to labels (reference: trackertraincode/eval.py).

Only `PerspectiveCorrector` is on the B200 path (the "perspective-correction terms" of the label
transforms); predictors, metrics and the ONNX wrappers of the reference's eval.py are out of scope.
"""
from __future__ import annotations

import math

import torch

from . import _native as N


class PerspectiveCorrector:
    """eval.py:485-544.  Same constructor and `corrected_rotation(image_sizes, coord, pose)` signature as the reference."""

    def __init__(self, fov):
        self._fov = fov
        self.f = 1.0 / math.tan(fov * math.pi / 180.0 * 0.5)

    def corrected_rotation(self, image_sizes: torch.Tensor, coord: torch.Tensor, pose: torch.Tensor) -> torch.Tensor:
        """image_sizes: [2] = (width, height) shared by all samples, or [B, 2]; coord [..., 3]; pose [..., 4] (xyzw).
        Returns the rotations transformed into the camera frame.  As in the reference both axes are divided by
        `half_image_size_tensor[0]` (eval.py:525): the half width for a [2] size, ROW 0 of a [B, 2] size."""
        out, _ = self._run(image_sizes, coord, pose, want_matrix=False)
        return out

    @staticmethod
    def _make_look_at_matrix(pos: torch.Tensor) -> torch.Tensor:
        """eval.py:531-544: rotation whose z axis points along `pos`, x axis kept horizontal.  pos [..., 3] (CUDA)."""
        if not pos.is_cuda:
            raise N.NativeError(f"tensor lives on {pos.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
        p = pos.to(torch.float32).reshape(-1, 3).contiguous()
        m = torch.empty((p.shape[0], 3, 3), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            N.check(N.lib.b200aug_corrected_rotation(None, 0, 1.0, 1.0, 0.0, p.data_ptr(), 3, None, None, m.data_ptr(), p.shape[0],
                                                     torch.cuda.current_stream(p.device).cuda_stream), "b200aug_corrected_rotation")
        return m.reshape(*pos.shape[:-1], 3, 3)

    def _run(self, image_sizes, coord, pose, want_matrix: bool):
        if not (coord.is_cuda and pose.is_cuda):
            raise N.NativeError("coord and pose must live on a CUDA device; this path has no CPU implementation")
        dev = coord.device
        lead = coord.shape[:-1]
        c = coord.to(torch.float32).reshape(-1, coord.shape[-1]).contiguous()
        q = pose.to(torch.float32).reshape(-1, 4).contiguous()
        n = c.shape[0]
        assert q.shape[0] == n, "coord and pose must have the same leading shape"
        half = 0.5 * torch.as_tensor(image_sizes).to(torch.float32)
        if half.dim() == 1:
            hs = half.cpu()
            div_x = div_y = float(hs[0])
            stride = 0
        else:
            half = half.reshape(-1, 2)
            assert half.shape[0] == n, "one image size per sample"
            h0 = half[0].cpu()
            div_x, div_y = float(h0[0]), float(h0[1])
            stride = 2
        half = half.to(dev).contiguous()
        out = torch.empty((n, 4), dtype=torch.float32, device=dev)
        mat = torch.empty((n, 3, 3), dtype=torch.float32, device=dev) if want_matrix else None
        with torch.cuda.device(dev):
            N.check(N.lib.b200aug_corrected_rotation(half.data_ptr(), stride, div_x, div_y, float(self.f), c.data_ptr(), c.shape[1],
                                                     q.data_ptr(), out.data_ptr(), mat.data_ptr() if want_matrix else None, n,
                                                     torch.cuda.current_stream(dev).cuda_stream), "b200aug_corrected_rotation")
        return out.reshape(*lead, 4), (mat.reshape(*lead, 3, 3) if want_matrix else None)
