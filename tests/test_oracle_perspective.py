"""The rotation-label oracle (oracle/perspective.py) against (i) outputs of the UNMODIFIED reference
(tests/golden/perspective.npz, made by tests/golden/make_golden_perspective.py) and (ii) the reference's own known-answer
tests for this path (test/test_eval.py:87-141)."""
import math
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle import perspective as P

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "perspective.npz"))
FOVS = (60.0, 90.0, 45.5)


@pytest.mark.parametrize("fov", FOVS)
def test_corrected_rotation_matches_reference(fov):
    assert abs(1.0 / math.tan(fov * math.pi / 180.0 * 0.5) - float(G[f"f_fov{fov:g}"])) < 1e-15
    per = np.stack([P.corrected_rotation(fov, G["image_sizes"][i], G["coord"][i], G["pose"][i]) for i in range(len(G["pose"]))])
    np.testing.assert_allclose(per, G[f"corrected_fov{fov:g}"], rtol=0, atol=5e-7)
    shared = P.corrected_rotation(fov, G["image_sizes"][0], G["coord"], G["pose"])
    np.testing.assert_allclose(shared, G[f"corrected_shared_fov{fov:g}"], rtol=0, atol=5e-7)


def test_matrix_conversions_match_reference():
    assert np.array_equal(P.tomatrix(G["pose"]), G["tomatrix"])
    np.testing.assert_allclose(P.from_matrix(G["tomatrix"]), G["from_matrix"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(P.from_matrix(G["special_mats"]), G["special_from_matrix"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(P.make_look_at_matrix(G["look_in"]), G["look_out"], rtol=0, atol=5e-7)
    # scipy as the truth, as test/test_torchquaternion.py does: from_matrix inverts tomatrix up to sign
    q = P.from_matrix(Rotation.from_quat(G["pose"]).as_matrix().astype(np.float32))
    assert (Rotation.from_quat(q) * Rotation.from_quat(G["pose"]).inv()).magnitude().max() < 1e-3


def fov_h(fov, aspect):  # test/test_eval.py:80-85
    return 2.0 * math.atan(1.0 / aspect * math.tan(fov / 2.0 * math.pi / 180.0)) * 180.0 / math.pi


KATS = [  # test/test_eval.py:87-127
    (90.0, [200, 100], [200.0, 50.0, 1.0], Rotation.identity(), Rotation.from_rotvec([0.0, 45.0, 0.0], degrees=True)),
    (90.0, [200, 100], [100.0, 100.0, 1.0], Rotation.identity(), Rotation.from_rotvec([-fov_h(90.0, 2.0) / 2.0, 0.0, 0.0], degrees=True)),
    (90.0, [200, 100], [100.0, 50.0, 1.0], Rotation.identity(), Rotation.identity()),
    (90.0, [200, 100], [100.0, 50.0, 1.0], Rotation.from_rotvec([10.0, 20.0, 30.0], degrees=True), Rotation.from_rotvec([10.0, 20.0, 30.0], degrees=True)),
]


@pytest.mark.parametrize("fov, image_size, coord, pose, expected", KATS)
def test_reference_perspective_kats(fov, image_size, coord, pose, expected):
    got = Rotation.from_quat(P.corrected_rotation(fov, np.array(image_size), np.array(coord, np.float32), pose.as_quat().astype(np.float32)))
    assert (expected.inv() * got).magnitude() * 180 / math.pi < 0.01


def test_reference_look_at_kats():  # test/test_eval.py:130-141
    np.testing.assert_allclose(P.make_look_at_matrix(np.array([0.0, 0.0, 1.0])), np.eye(3))
    m = P.make_look_at_matrix(np.array([1.0, 1.0, 1.0]))
    np.testing.assert_allclose(m[:, 2], np.full(3, 1.0 / math.sqrt(3.0)), rtol=1e-6)
    assert abs(np.dot(m[:, 0], [0.0, 1.0, 0.0])) < 1e-6 and m[0, 0] > 0.1 and m[1, 1] > 0.1
