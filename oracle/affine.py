"""2x3 similarity algebra in float32 with an explicit operation order -- TEST INFRASTRUCTURE.

Follows trackertraincode/neuralnets/affine2d.py (Affine2d) and trackertraincode/neuralnets/math.py:8-20.
Matrices are numpy float32 arrays of shape [..., 2, 3].  Every arithmetic step is a separate float32 numpy
ufunc (correctly rounded, never contracted), in the order the reference's torch ops execute, so the CUDA
kernels can reproduce the bits with __fmul_rn/__fadd_rn.

Rounding facts pinned against torch 2.11 CPU in the authoring container (tests/golden/make_golden.py):
  * `torch.matmul(2x2, 2x2, out=)` (Affine2d.__matmul__, affine2d.py:177) evaluates c = fma(a1, b1, a0*b0);
  * `matvecmul` (math.py:8-14) evaluates (a0*b0) + (a1*b1) unfused;
  * `torch.norm` over the 2x2 block (affine2d.py:189-192) is sqrt of the sequential float32 sum of squares;
  * `torch.cos/sin` agree with the correctly rounded value for the production angles (0, +-30 deg, +-90 deg)
    but differ by 1 ulp for ~5 % of arbitrary angles; the oracle uses the correctly rounded value.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
SQRT2 = np.sqrt(F32(2.0))  # affine2d.py:7


def _f(x):
    return np.asarray(x, dtype=F32)


def fma32(a, b, c):
    """float32 fma emulated in float64 (the product is exact there; double rounding is a ~2^-29 event)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def cos32(a):
    return np.cos(np.asarray(a, np.float64)).astype(F32)


def sin32(a):
    return np.sin(np.asarray(a, np.float64)).astype(F32)


def identity(shape=()):
    m = np.zeros(tuple(shape) + (2, 3), F32)
    m[..., 0, 0] = 1
    m[..., 1, 1] = 1
    return m


def trs(translations=None, angles=None, scales=None, cs_sn=None):
    """Affine2d.trs (affine2d.py:37-59).  `cs_sn=(cos, sin)` overrides the trig evaluation (a caller that has the
    reference's own `torch.cos/sin(angles)` values passes them to stay bit-exact for arbitrary angles)."""
    if translations is not None:
        translations = _f(translations)
        shape = translations.shape[:-1]
    elif angles is not None or cs_sn is not None:
        shape = _f(angles if angles is not None else cs_sn[0]).shape
    else:
        shape = _f(scales).shape
    m = np.zeros(shape + (2, 3), F32)
    if angles is None and cs_sn is None:
        m[..., 0, 0] = 1
        m[..., 1, 1] = 1
        if scales is not None:
            m[..., :, :2] = m[..., :, :2] * _f(scales)[..., None, None]
    else:
        cs, sn = (cos32(angles), sin32(angles)) if cs_sn is None else (_f(cs_sn[0]), _f(cs_sn[1]))
        if scales is not None:
            cs = cs * _f(scales)
            sn = sn * _f(scales)
        m[..., 0, 0] = cs
        m[..., 0, 1] = -sn
        m[..., 1, 0] = sn
        m[..., 1, 1] = cs
    if translations is not None:
        m[..., :, 2] = translations
    return m


def range_remap_2d(inmin, inmax, outmin, outmax):
    """Affine2d.range_remap_2d (affine2d.py:118-133): per-axis scale + offset."""
    inmin, inmax, outmin, outmax = np.broadcast_arrays(_f(inmin), _f(inmax), _f(outmin), _f(outmax))
    s = (outmax - outmin) / (inmax - inmin)
    m = np.zeros(inmin.shape[:-1] + (2, 3), F32)
    m[..., 0, 0] = s[..., 0]
    m[..., 1, 1] = s[..., 1]
    m[..., :, 2] = outmin - inmin * s
    return m


def position_normalization(w, h):
    """affinetrafo.py:11-12: pixels [0,w]x[0,h] -> [-1,1]^2."""
    return range_remap_2d([0.0, 0.0], [w, h], [-1.0, -1.0], [1.0, 1.0])


def position_unnormalization(w, h):
    """affinetrafo.py:15-16."""
    return range_remap_2d([-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [w, h])


def matvec2(R, v):
    """math.py:8-14 matvecmul for a 2x2 block: unfused two-term dot products."""
    R, v = _f(R), _f(v)
    x = R[..., 0, 0] * v[..., 0] + R[..., 0, 1] * v[..., 1]
    y = R[..., 1, 0] * v[..., 0] + R[..., 1, 1] * v[..., 1]
    return np.stack([x, y], axis=-1)


def affinevecmul(m, v):
    """math.py:17-20: o = R v; o += t."""
    m = _f(m)
    return matvec2(m[..., :, :2], v) + m[..., :, 2]


def compose(a, b):
    """Affine2d.__matmul__ (affine2d.py:173-180): R = Ra Rb (fma on the second term), T = Ra Tb + Ta."""
    a, b = np.broadcast_arrays(_f(a), _f(b))
    m = np.empty(a.shape, F32)
    for i in range(2):
        for j in range(2):
            m[..., i, j] = fma32(a[..., i, 1], b[..., 1, j], a[..., i, 0] * b[..., 0, j])
    m[..., :, 2] = matvec2(a[..., :, :2], b[..., :, 2]) + a[..., :, 2]
    return m


def det(m):
    """affine2d.py:195-197."""
    m = _f(m)
    return m[..., 0, 0] * m[..., 1, 1] - m[..., 0, 1] * m[..., 1, 0]


def scales(m):
    """affine2d.py:189-192: Frobenius norm of the 2x2 block / sqrt(2), float32 sequential sum."""
    m = _f(m)
    ss = ((m[..., 0, 0] * m[..., 0, 0] + m[..., 0, 1] * m[..., 0, 1]) + m[..., 1, 0] * m[..., 1, 0]) + m[..., 1, 1] * m[..., 1, 1]
    return np.sqrt(ss) / SQRT2


def inv(m):
    """Affine2d.inv (affine2d.py:182-186).  The reference calls LAPACK (`torch.inverse`); the closed form below
    agrees to float32 round-off, which is all the label tolerance (1e-4 rel) needs."""
    m = _f(m)
    d = det(m)
    r = np.empty(m.shape, F32)
    r[..., 0, 0] = m[..., 1, 1] / d
    r[..., 0, 1] = -m[..., 0, 1] / d
    r[..., 1, 0] = -m[..., 1, 0] / d
    r[..., 1, 1] = m[..., 0, 0] / d
    r[..., :, 2] = -matvec2(r[..., :, :2], m[..., :, 2])
    return r
