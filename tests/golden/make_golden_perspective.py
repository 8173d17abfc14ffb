#!/usr/bin/env python
"""Generate tests/golden/perspective.npz by running the UNMODIFIED reference PerspectiveCorrector (eval.py:485-544),
torchquaternion.from_matrix / mult / tomatrix (neuralnets/torchquaternion.py) and unnormalize-side helpers on seeded inputs.

Run in the authoring container only:  python tests/golden/make_golden_perspective.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", HERE]

import kornia_stub  # noqa: E402

kornia_stub.install()

import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.spatial.transform import Rotation  # noqa: E402

from trackertraincode.eval import PerspectiveCorrector  # noqa: E402
from trackertraincode.neuralnets import torchquaternion as tq  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    n = 256
    sizes = np.stack([rng.choice([640.0, 450.0, 1280.0, 200.0], n), rng.choice([480.0, 450.0, 720.0, 100.0], n)], -1).astype(np.float32)
    coord = np.concatenate([rng.uniform(-0.1, 1.1, (n, 2)) * sizes, rng.uniform(20.0, 200.0, (n, 1))], -1).astype(np.float32)
    coord[:4, :2] = 0.5 * sizes[:4]                      # dead centre: look-at is the identity
    coord[4:8, 0] = sizes[4:8, 0]                        # on the right edge
    pose = Rotation.random(n, random_state=7).as_quat().astype(np.float32)
    out = {"image_sizes": sizes, "coord": coord, "pose": pose}
    for fov in (60.0, 90.0, 45.5):
        pc = PerspectiveCorrector(fov)
        # the reference divides by half_image_size_tensor[0] (eval.py:525): with a [B,2] tensor that is ROW 0 of the batch;
        # the per-sample call below is the semantics of the reference's own test (one image size per call)
        per_sample = torch.stack([pc.corrected_rotation(torch.from_numpy(sizes[i]), torch.from_numpy(coord[i]), torch.from_numpy(pose[i]))
                                  for i in range(n)])
        out[f"corrected_fov{fov:g}"] = per_sample.numpy()
        # one shared image size for the whole batch (how eval scripts call it): rows broadcast
        shared = pc.corrected_rotation(torch.from_numpy(sizes[0]), torch.from_numpy(coord), torch.from_numpy(pose))
        out[f"corrected_shared_fov{fov:g}"] = shared.numpy()
        out[f"f_fov{fov:g}"] = np.float64(pc.f)
    mats = tq.tomatrix(torch.from_numpy(pose))
    out["tomatrix"] = mats.numpy()
    out["from_matrix"] = tq.from_matrix(mats).numpy()
    # the four branches of from_matrix: rotations by ~180 degrees about x, y, z and small ones
    special = Rotation.from_rotvec(np.array([[3.1, 0, 0], [0, 3.1, 0], [0, 0, 3.1], [0.01, 0.02, 0.03], [2.0, 2.0, 0.5]])).as_matrix().astype(np.float32)
    out["special_mats"] = special
    out["special_from_matrix"] = tq.from_matrix(torch.from_numpy(special)).numpy()
    pos = rng.normal(size=(64, 3)).astype(np.float32) + np.float32([0, 0, 2.0])
    out["look_in"] = pos
    out["look_out"] = PerspectiveCorrector._make_look_at_matrix(torch.from_numpy(pos)).numpy()
    np.savez_compressed(os.path.join(HERE, "perspective.npz"), **out)
    print("wrote perspective.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
