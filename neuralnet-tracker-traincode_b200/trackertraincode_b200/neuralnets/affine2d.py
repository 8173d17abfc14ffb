"""`Affine2d`: N x 2 x 3 similarity transforms as a thin torch wrapper (host-side plumbing, not the hot path).

Interface of trackertraincode/neuralnets/affine2d.py:12-223 (trs, range_remap_2d, @, inv, scales, det, ...), kept so
that callers such as eval.py (`Affine2d(batch["image_backtransform"])`) and `apply_affine2d(trafo, ...)` work unchanged.
The per-sample transforms of the fused path are composed inside the CUDA kernel, not here.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

_SQRT2 = math.sqrt(2.0)


def _rot_block(angles: torch.Tensor, scales: Optional[torch.Tensor]):
    cs, sn = torch.cos(angles), torch.sin(angles)
    if scales is not None:
        cs, sn = cs * scales, sn * scales
    return torch.stack([torch.stack([cs, -sn], -1), torch.stack([sn, cs], -1)], -2)


class Affine2d:
    def __init__(self, m: torch.Tensor):
        assert m.dim() >= 2 and tuple(m.shape[-2:]) == (2, 3), f"bad shape {tuple(m.shape)}"
        self.m = m.detach().to(torch.float32)

    # constructors ------------------------------------------------------------------------------------------
    @staticmethod
    def identity(device=None) -> "Affine2d":
        return Affine2d(torch.eye(2, 3, device=device))

    @staticmethod
    def trs(translations=None, angles=None, scales=None) -> "Affine2d":
        ref = next(x for x in (translations, angles, scales) if x is not None)
        shape = ref.shape[:-1] if ref is translations else ref.shape
        if angles is not None:
            R = _rot_block(angles.to(torch.float32), scales)
        else:
            R = torch.eye(2, device=ref.device).expand(*shape, 2, 2)
            if scales is not None:
                R = R * scales[..., None, None]
        T = translations if translations is not None else torch.zeros(*shape, 2, device=ref.device)
        return Affine2d(torch.cat([R.expand(*shape, 2, 2), T[..., None].to(torch.float32)], dim=-1))

    @staticmethod
    def trs_inv(translations=None, angles=None, scales=None) -> "Affine2d":
        return Affine2d.trs(translations, angles, scales).inv()

    @staticmethod
    def horizontal_flip(xcenter: torch.Tensor) -> "Affine2d":
        z, o = torch.zeros_like(xcenter), torch.ones_like(xcenter)
        return Affine2d(torch.stack([torch.stack([-o, z, 2 * xcenter], -1), torch.stack([z, o, z], -1)], -2))

    @staticmethod
    def range_remap_2d(inmin, inmax, outmin, outmax) -> "Affine2d":
        inmin, inmax, outmin, outmax = (torch.as_tensor(x, dtype=torch.float32) for x in (inmin, inmax, outmin, outmax))
        s = (outmax - outmin) / (inmax - inmin)
        t = outmin - inmin * s
        z = torch.zeros_like(s[..., 0])
        return Affine2d(torch.stack([torch.stack([s[..., 0], z, t[..., 0]], -1), torch.stack([z, s[..., 1], t[..., 1]], -1)], -2))

    @staticmethod
    def range_remap(inmin, inmax, outmin, outmax) -> "Affine2d":
        inmin, inmax, outmin, outmax = (torch.as_tensor(x, dtype=torch.float32) for x in (inmin, inmax, outmin, outmax))
        s = (outmax - outmin) / (inmax - inmin)
        t = outmin[..., None] - inmin[..., None] * s[..., None] if outmin.dim() == s.dim() else outmin - inmin * s[..., None]
        t = torch.broadcast_to(t, s.shape + (2,))
        z = torch.zeros_like(s)
        return Affine2d(torch.stack([torch.stack([s, z, t[..., 0]], -1), torch.stack([z, s, t[..., 1]], -1)], -2))

    # accessors ---------------------------------------------------------------------------------------------
    def tensor(self) -> torch.Tensor:
        return self.m

    def tensor33(self) -> torch.Tensor:
        bottom = self.m.new_tensor([0.0, 0.0, 1.0]).expand(*self.m.shape[:-2], 1, 3)
        return torch.cat([self.m, bottom], dim=-2)

    def to(self, *a, **k) -> "Affine2d":
        return Affine2d(self.m.to(*a, **k))

    @property
    def R(self):
        return self.m[..., :2, :2]

    @property
    def T(self):
        return self.m[..., :2, 2]

    @property
    def R33(self):
        r = torch.zeros(*self.m.shape[:-2], 3, 3, device=self.m.device)
        r[..., :2, :2] = self.R
        r[..., 2, 2] = 1.0
        return r

    @property
    def shape(self):
        return self.m.shape[:-2]

    def size(self, i):
        return self.m.size(i)

    @property
    def scales(self):
        return torch.linalg.matrix_norm(self.R) / _SQRT2

    @property
    def det(self):
        return self.m[..., 0, 0] * self.m[..., 1, 1] - self.m[..., 0, 1] * self.m[..., 1, 0]

    # algebra -----------------------------------------------------------------------------------------------
    def __matmul__(self, other: "Affine2d") -> "Affine2d":
        a, b = torch.broadcast_tensors(self.m, other.m)
        R = a[..., :2, :2] @ b[..., :2, :2]
        T = (a[..., :2, :2] @ b[..., :2, 2:3])[..., 0] + a[..., :2, 2]
        return Affine2d(torch.cat([R, T[..., None]], dim=-1))

    def inv(self) -> "Affine2d":
        Ri = torch.linalg.inv(self.R)
        Ti = -(Ri @ self.T[..., None])[..., 0]
        return Affine2d(torch.cat([Ri, Ti[..., None]], dim=-1))

    def __getitem__(self, idx):
        return Affine2d(self.m[idx])

    def reshape(self, shape):
        return Affine2d(self.m.reshape(tuple(shape) + (2, 3)))

    def view(self, *shape):
        return Affine2d(self.m.view(*shape, 2, 3))

    def expand(self, *shape):
        return Affine2d(self.m.expand(*shape, -1, -1))

    def repeat(self, size):
        return Affine2d(self.m.repeat(tuple(size) + (1, 1)))


def roi_normalizing_transform(roi: torch.Tensor) -> Affine2d:
    """affine2d.py:214-223: map each roi to [-1, 1]^2."""
    assert roi.shape[-1] == 4
    lo = roi.new_full(roi.shape[:-1] + (2,), -1.0)
    return Affine2d.range_remap_2d(roi[..., :2], roi[..., 2:], lo, -lo)
