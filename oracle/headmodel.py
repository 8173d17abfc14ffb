"""PutRoiFromLandmarks(extend_to_forehead=True): the roi from the posed face-model vertices -- TEST INFRASTRUCTURE.

Restates, in numpy:
  trackertraincode/datatransformation/batch/misc.py:14-26        _create_roi
  trackertraincode/neuralnets/modelcomponents.py:38-56, 85-94    rigid_transformation_25d, PosedDeformableHead.forward
  trackertraincode/facemodel/bfm.py:49-96                        the scaled mean shape / deformation bases, ScaledBfmModule
  trackertraincode/neuralnets/torchquaternion.py:51-67           rotate (q p q^-1)

Pinned against outputs of the unmodified reference (tests/golden/headroi.npz, tests/golden/make_golden_headroi.py): the
real face model for the roi, the reference's PosedDeformableHead on a synthetic model for non-zero shape parameters.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def scaled_model(u, w_shp, w_exp, shape_dim=40, exp_dim=10):
    """bfm.py:24-72: (vertices [V,3], deform_base [shape_dim + exp_dim, V, 3]) from the raw arrays of bfm_noneck_v3.pkl."""
    u = np.asarray(u).astype(F32)
    w_shp = np.asarray(w_shp).astype(F32)[..., :shape_dim]
    w_exp = np.asarray(w_exp).astype(F32)[..., :exp_dim]
    V = u.shape[0] // 3
    flip = np.array([[[1.0, -1.0, -1.0]]])
    shp = (20.0 * w_shp.reshape((V, 3, -1))).transpose([2, 0, 1]) * flip
    exp = (5.0e-5 * w_exp.reshape((V, 3, -1))).transpose([2, 0, 1]) * flip
    vertices = u.reshape((-1, 3)) * 1.0e-5 * np.array([[1.0, -1.0, -1.0]], dtype="f4")
    vertices = vertices - np.array([0.0, -0.26, -0.9], dtype="f4")[None, :]
    return np.ascontiguousarray(vertices, F32), np.concatenate([shp, exp], 0).astype(F32)


def quat_rotate(q, p):
    """torchquaternion.rotate: q p q^-1 for unit quaternions (i, j, k, w), float32."""
    q = np.asarray(q, F32)
    p = np.asarray(p, F32)
    x, y, z, w = (q[..., None, k] for k in range(4))
    px, py, pz = p[..., 0], p[..., 1], p[..., 2]
    # t = q * (p, 0)
    tw = -x * px - y * py - z * pz
    tx = w * px + y * pz - z * py
    ty = w * py + z * px - x * pz
    tz = w * pz + x * py - y * px
    # t * conj(q)
    ox = -tw * x + tx * w - ty * z + tz * y
    oy = -tw * y + ty * w - tz * x + tx * z
    oz = -tw * z + tz * w - tx * y + ty * x
    return np.stack([ox, oy, oz], -1).astype(F32)


def posed_vertices(vertices, deform_base, coord, quat, shapeparams=None):
    """PosedDeformableHead.forward: [..., V, 3]."""
    verts = np.asarray(vertices, F32)
    if shapeparams is not None:
        sp = np.asarray(shapeparams, F32)
        verts = np.sum(np.asarray(deform_base, F32) * sp[..., None, None], axis=-3, dtype=F32) + verts
    coord = np.asarray(coord, F32)
    pos = quat_rotate(quat, verts) * coord[..., None, 2:]
    pos[..., :2] += coord[..., None, :2]
    return pos.astype(F32)


def head_roi(vertices, deform_base, coord, quat, shapeparams=None):
    """misc.py:18-25 with extend_to_forehead: [min_x, min_y, max_x, max_y] over all posed vertices."""
    v = posed_vertices(vertices, deform_base, coord, quat, shapeparams)
    return np.concatenate([v[..., :2].min(-2), v[..., :2].max(-2)], -1).astype(F32)
