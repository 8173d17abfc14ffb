#!/usr/bin/env python
"""BASELINE.json configs[0]: the reference's own augmentation chain on its bundled fixture `aflw2kmini.h5`
(16 x 450x450 JPEG frames + rois / coords / quats / pt3d_68 / shapeparams), batch 64 = the 16 samples under 4 parameter
draws, 129 x 129 gray crops.  Runs the UNMODIFIED reference (make_golden.run_case) in the authoring container and writes
tests/golden/aflw2kmini.npz: the JPEG blobs as stored in the file, the labels, the draws and the reference's outputs.
The file is read with tests/golden/minihdf5.py (no h5py in this image).   python tests/golden/make_golden_aflw2kmini.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import cv2  # noqa: E402
import numpy as np  # noqa: E402

import make_golden  # noqa: E402  (sets up the stubs and imports the reference)
import minihdf5  # noqa: E402

DRAWS = 4
S = 129


def main():
    f = minihdf5.File("/root/reference/aflw2kmini.h5")
    blobs = f.read("images")
    labels = dict(roi=f.read("rois"), coord=f.read("coords"), pose=f.read("quats"), pt3d_68=f.read("pt3d_68"), shapeparam=f.read("shapeparams"))
    n = len(blobs)
    frames = [cv2.imdecode(b, 0) for b in blobs]  # datasets/preprocessing.py:42-54
    assert all(fr.shape == (450, 450) for fr in frames)
    rng = np.random.default_rng(20240)
    rows, draws = [], []
    for k in range(DRAWS):
        for i in range(n):
            # the sampler of geometric.py:58-84 (ext 1.1, raug 30), seeded; eval-style draw for k == 0
            scale = np.float32(1.1) if k == 0 else np.float32(np.clip(0.1 * rng.normal(), -0.5, 0.5) + 1.1)
            tr = np.zeros(2, np.float32) if k == 0 else np.clip(0.5 * rng.normal(size=2), -1, 1).astype(np.float32)
            angle = np.float32(0.0) if k == 0 else np.float32(rng.choice([0.0, 0.0, 0.0, 0.0, 30.0, -30.0]) * np.pi / 180.0)
            c = dict(wh=(450, 450), image=frames[i], out_size=S, scale=scale, angle=angle, translation=tr,
                     do_flip=bool(k and rng.integers(0, 2)), rot_dir=int(rng.choice([-1, 0, 0, 0, 0, 0, 0, 1])) if k == 3 else 0,
                     **{key: v[i] for key, v in labels.items()})
            rows.append(make_golden.run_case(c))
            draws.append((scale, angle, tr, c["do_flip"], c["rot_dir"]))
    out = {k: np.stack([r[k] for r in rows], 0) for k in ("view_roi", "tr", "flip_image", "final_roi", "final_coord", "final_pose", "final_pt3d_68",
                                                         "final_shapeparam")}
    out["jpeg_bytes"] = np.concatenate(blobs)
    out["jpeg_offsets"] = np.cumsum([0] + [len(b) for b in blobs]).astype(np.int64)
    for key, v in labels.items():
        out["in_" + key] = v
    out["scales"] = np.array([d[0] for d in draws], np.float32)
    out["angles"] = np.array([d[1] for d in draws], np.float32)
    out["translations"] = np.stack([d[2] for d in draws]).astype(np.float32)
    out["do_flip"] = np.array([d[3] for d in draws], bool)
    out["rot_dir"] = np.array([d[4] for d in draws], np.int8)
    path = os.path.join(HERE, "aflw2kmini.npz")
    np.savez_compressed(path, **out)
    print("aflw2kmini.npz:", {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
