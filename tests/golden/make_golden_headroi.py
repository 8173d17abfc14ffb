#!/usr/bin/env python
"""Generate tests/golden/headroi.npz: PutRoiFromLandmarks(extend_to_forehead=True) of the UNMODIFIED reference
(trackertraincode/datatransformation/batch/misc.py:9-31) with its real face model.

Run in the authoring container only:  python tests/golden/make_golden_headroi.py

The face model (24 MB) cannot travel, so the file holds:
  coord, pose, pt3d_68   seeded poses (AFLW2k-like ranges)
  roi_full               the reference's roi for them (full model; the reference reads the shape parameters only when the
                         sample has a key "shapeparams", which no dataset has -- so this is the mean shape)
  hull_vertices          the few hundred mean-shape vertices that are extreme in x or y under some rotation (the roi only
                         depends on those), checked here to give exactly roi_full through the reference's own module
  syn_*                  a synthetic deformable model (random vertices / bases / shape parameters) through the reference's
                         PosedDeformableHead, for the path with non-zero shape parameters
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", HERE]

import kornia_stub  # noqa: E402

kornia_stub.install()

import numpy as np  # noqa: E402
import torch  # noqa: E402

import trackertraincode.datatransformation as dtr  # noqa: E402
from trackertraincode.datasets.batch import Batch, Metadata  # noqa: E402
from trackertraincode.datasets.dshdf5pose import FieldCategory  # noqa: E402
from trackertraincode.neuralnets.modelcomponents import PosedDeformableHead  # noqa: E402
from trackertraincode.neuralnets.rotrepr import QuatRepr  # noqa: E402

N = 48


def poses(rng, n):
    q = rng.standard_normal((n, 4))
    q[:, 3] += 2.5  # mostly frontal, some strongly turned
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    coord = np.stack([rng.uniform(100, 350, n), rng.uniform(100, 350, n), rng.uniform(30, 120, n)], 1)
    return coord.astype(np.float32), q.astype(np.float32)


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(77)
    coord, pose = poses(rng, N)
    pts = rng.uniform(80, 370, (N, 68, 3)).astype(np.float32)
    put = dtr.batch.PutRoiFromLandmarks(extend_to_forehead=True)
    cats = {"coord": FieldCategory.xys, "pose": FieldCategory.quat, "pt3d_68": FieldCategory.points, "shapeparam": FieldCategory.general}
    rois = []
    for i in range(N):
        s = Batch(Metadata((450, 450), 0, "golden", None, categories=dict(cats)),
                  {"coord": torch.from_numpy(coord[i]), "pose": torch.from_numpy(pose[i]), "pt3d_68": torch.from_numpy(pts[i]),
                   "shapeparam": torch.from_numpy(rng.standard_normal(50).astype(np.float32))})
        rois.append(put(s)["roi"].numpy())
    roi_full = np.stack(rois)

    # the vertices that can be extreme: argmin / argmax of x and y of the rotated mean shape over many rotations
    model = put.headmodel.deformable_head
    verts = model.vertices.numpy()
    keep = set()
    qs = rng.standard_normal((6000, 4))
    qs = np.concatenate([qs / np.linalg.norm(qs, axis=1, keepdims=True), pose.astype(np.float64)], 0)
    for q in qs:
        x, y, z, w = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]])
        p = verts.astype(np.float64) @ R.T
        for col in (0, 1):
            order = np.argsort(p[:, col])
            keep.update(order[:3].tolist())
            keep.update(order[-3:].tolist())
    idx = np.asarray(sorted(keep))
    hull = verts[idx]
    # check: the reference's module on the reduced vertex set gives the same roi, bit for bit
    model.vertices = torch.from_numpy(hull.copy())
    model.deform_base = model.deform_base[:, idx, :]
    for i in range(N):
        v = put.headmodel(torch.from_numpy(coord[i]), QuatRepr(torch.from_numpy(pose[i])), torch.zeros(50))
        r = torch.cat([torch.amin(v[..., :2], dim=-2), torch.amax(v[..., :2], dim=-2)], 0).numpy()
        assert np.array_equal(r, roi_full[i]), i

    # synthetic deformable model through the reference's PosedDeformableHead
    class Syn(torch.nn.Module):
        def __init__(self, vertices, base):
            super().__init__()
            self.vertices, self.deform_base = vertices, base

        def forward(self, shapeparams):
            verts = self.deform_base * shapeparams[..., None, None]
            verts = torch.sum(verts, dim=-3)
            verts += self.vertices
            return verts

    V, K = 400, 50
    syn_vertices = rng.standard_normal((V, 3)).astype(np.float32)
    syn_base = (0.05 * rng.standard_normal((K, V, 3))).astype(np.float32)
    syn_shape = rng.standard_normal((N, K)).astype(np.float32)
    # (the forward of Syn is ScaledBfmModule.forward, bfm.py:91-95, which needs the BFM pickle to construct)
    head = PosedDeformableHead(Syn(torch.from_numpy(syn_vertices), torch.from_numpy(syn_base)))
    syn_roi = []
    for i in range(N):
        v = head(torch.from_numpy(coord[i]), QuatRepr(torch.from_numpy(pose[i])), torch.from_numpy(syn_shape[i]))
        syn_roi.append(torch.cat([torch.amin(v[..., :2], dim=-2), torch.amax(v[..., :2], dim=-2)], 0).numpy())
    np.savez_compressed(os.path.join(HERE, "headroi.npz"), coord=coord, pose=pose, pt3d_68=pts, roi_full=roi_full, hull_vertices=hull,
                        syn_vertices=syn_vertices, syn_base=syn_base, syn_shape=syn_shape, syn_roi=np.stack(syn_roi))
    print("hull vertices:", hull.shape, "of", verts.shape, "roi_full[0]", roi_full[0])


if __name__ == "__main__":
    main()
