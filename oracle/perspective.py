"""CPU restatement (numpy float32) of the rotation post-processing on the label path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package; the product path never does.

Follows the reference:
  PerspectiveCorrector.corrected_rotation   trackertraincode/eval.py:491-529
  PerspectiveCorrector._make_look_at_matrix trackertraincode/eval.py:531-544
  torchquaternion.mult / tomatrix / from_matrix / positivereal   trackertraincode/neuralnets/torchquaternion.py:23-48, 70-91, 94-168, 216-218
Pinned by tests/golden/perspective.npz (outputs of the unmodified reference, tests/golden/make_golden_perspective.py) and by
the reference's own known-answer tests (test/test_eval.py:87-141), see tests/test_oracle_perspective.py.
"""
import math

import numpy as np

F = np.float32


def quat_mult(u, v):
    """torchquaternion.py:40-48 (xyzw)."""
    u, v = np.asarray(u, F), np.asarray(v, F)
    ui, uj, uk, uw = (u[..., k] for k in range(4))
    vi, vj, vk, vw = (v[..., k] for k in range(4))
    w = uw * vw - ui * vi - uj * vj - uk * vk
    i = ui * vw + uw * vi - uk * vj + uj * vk
    j = uj * vw + uk * vi + uw * vj - ui * vk
    k = uk * vw - uj * vi + ui * vj + uw * vk
    return np.stack([i, j, k, w], -1).astype(F)


def tomatrix(q):
    """torchquaternion.py:70-91."""
    q = np.asarray(q, F)
    qi, qj, qk, qw = (q[..., k] for k in range(4))
    o = np.empty(q.shape[:-1] + (3, 3), F)
    two, one = F(2), F(1)
    o[..., 0, 0] = one - two * (qj * qj + qk * qk)
    o[..., 1, 0] = two * (qi * qj + qk * qw)
    o[..., 2, 0] = two * (qi * qk - qj * qw)
    o[..., 0, 1] = two * (qi * qj - qk * qw)
    o[..., 1, 1] = one - two * (qi * qi + qk * qk)
    o[..., 2, 1] = two * (qj * qk + qi * qw)
    o[..., 0, 2] = two * (qi * qk + qj * qw)
    o[..., 1, 2] = two * (qj * qk - qi * qw)
    o[..., 2, 2] = one - two * (qi * qi + qj * qj)
    return o


def from_matrix(m):
    """torchquaternion.py:94-168: four candidates, argmax of the clamped square-root arguments, positivereal()."""
    m = np.asarray(m, F)
    shape = m.shape[:-2]
    m = m.reshape(-1, 3, 3)
    d0, d1, d2 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    args = np.stack([-d0 - d1 + d2, -d0 + d1 - d2, d0 - d1 - d2, d0 + d1 + d2], -1).astype(F) + F(1)  # k, j, i, w
    args = np.maximum(args, F(1e-6))
    qx = (np.sqrt(args) * F(0.5)).astype(F)
    pick = np.argmax(args, -1)
    out = np.empty((m.shape[0], 4), F)
    q25 = F(0.25)
    for n in range(m.shape[0]):
        a, d = m[n], qx[n, pick[n]]
        v = lambda x, y: q25 * (x + y) / d  # noqa: E731
        if pick[n] == 0:
            qw, qi, qj, qk = v(a[1, 0], -a[0, 1]), v(a[2, 0], a[0, 2]), v(a[1, 2], a[2, 1]), d
        elif pick[n] == 1:
            qw, qi, qk, qj = v(a[0, 2], -a[2, 0]), v(a[1, 0], a[0, 1]), v(a[1, 2], a[2, 1]), d
        elif pick[n] == 2:
            qw, qj, qk, qi = v(a[2, 1], -a[1, 2]), v(a[1, 0], a[0, 1]), v(a[0, 2], a[2, 0]), d
        else:
            qi, qj, qk, qw = v(a[2, 1], -a[1, 2]), v(a[0, 2], -a[2, 0]), v(a[1, 0], -a[0, 1]), d
        out[n] = np.array([qi, qj, qk, qw], F) * np.sign(qw).astype(F)
    return out.reshape(*shape, 4)


def make_look_at_matrix(pos):
    """eval.py:531-544 (note `y / |x|`, :542)."""
    pos = np.asarray(pos, F)
    z = pos / np.linalg.norm(pos, axis=-1, keepdims=True).astype(F)
    up = np.broadcast_to(np.array([0, 1, 0], F), z.shape)
    x = np.cross(up, z).astype(F)
    x = x / np.linalg.norm(x, axis=-1, keepdims=True).astype(F)
    y = np.cross(z, x).astype(F)
    y = y / np.linalg.norm(x, axis=-1, keepdims=True).astype(F)
    return np.stack([x, y, z], -1).astype(F)


def corrected_rotation(fov, image_sizes, coord, pose):
    """eval.py:491-529.  image_sizes [2] or [B,2]; both axes are divided by half_image_size_tensor[0] as in the reference."""
    f = 1.0 / math.tan(fov * math.pi / 180.0 * 0.5)
    coord, pose = np.asarray(coord, F), np.asarray(pose, F)
    half = (F(0.5) * np.asarray(image_sizes).astype(F)).astype(F)
    xy = (coord[..., :2] - half) / half[0]
    xyz = np.concatenate([xy, np.full(xy.shape[:-1] + (1,), F(f), F)], -1).astype(F)
    return quat_mult(from_matrix(make_look_at_matrix(xyz)), pose)
