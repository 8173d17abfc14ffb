"""PutRoiFromLandmarks(extend_to_forehead=True): the numpy restatement (oracle/headmodel.py) against outputs of the
unmodified reference with its real face model (tests/golden/headroi.npz, made by tests/golden/make_golden_headroi.py)."""
import os

import numpy as np
import pytest

from oracle import headmodel as hm

GOLD = os.path.join(os.path.dirname(__file__), "golden", "headroi.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_roi_of_the_mean_shape_matches_reference(gold):
    r = hm.head_roi(gold["hull_vertices"], None, gold["coord"], gold["pose"])
    np.testing.assert_allclose(r, gold["roi_full"], rtol=1e-6, atol=2e-4)
    assert (r[:, 2:] > r[:, :2]).all()


def test_deformed_model_matches_reference_module(gold):
    r = hm.head_roi(gold["syn_vertices"], gold["syn_base"], gold["coord"], gold["pose"], gold["syn_shape"])
    np.testing.assert_allclose(r, gold["syn_roi"], rtol=1e-6, atol=2e-4)
    # the shape parameters matter (the test would pass trivially otherwise)
    r0 = hm.head_roi(gold["syn_vertices"], gold["syn_base"], gold["coord"], gold["pose"])
    assert np.abs(r0 - gold["syn_roi"]).max() > 1.0


def test_quat_rotate_is_a_rotation():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((5, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    p = rng.standard_normal((5, 7, 3)).astype(np.float32)
    out = hm.quat_rotate(q, p)
    np.testing.assert_allclose(np.linalg.norm(out, axis=-1), np.linalg.norm(p, axis=-1), rtol=1e-5)
    ident = hm.quat_rotate(np.float32([[0, 0, 0, 1]]), p[:1])
    np.testing.assert_allclose(ident, p[:1], atol=1e-7)
