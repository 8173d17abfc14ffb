"""The oracle's evaluation-side chain (focus with insert_backtransform -> normalize -> unnormalize -> back-transform, and
the `image_backtransform` rewriting of apply_affine2d) against golden vectors produced by the UNMODIFIED reference's
Predictor.predict_batch (tests/golden/make_golden_backtransform.py; eval.py:149-212, tensors/affinetrafo.py:137-147)."""
import os

import numpy as np
import pytest

import cases
from oracle import affine, geometric as ogeo, normalization as onrm
from oracle.geometric import RoiFocusParams, Sample
from oracle.labels import apply_affine2d

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "backtransform.npz"))
S = 129
CATS = dict(image="img", roi="roi")
PRED_CATS = dict(coord="xys", pose="q", pt3d_68="pts")


def focus(c, size=S, ext=1.1, insert=True, sample=None):
    if sample is None:
        sample = Sample(c["wh"], {"image": c["image"][..., None], "roi": c["roi"]}, dict(CATS))
    return ogeo.focus_roi(sample, RoiFocusParams(ext, 0.0, (0.0, 0.0)), size, insert_backtransform=insert)[0]


@pytest.mark.parametrize("j", range(len(GOLD["cases"])))
def test_predictor_chain(j):
    c = cases.make_case(int(GOLD["cases"][j]))
    s = focus(c)
    np.testing.assert_allclose(s.data["image_backtransform"], GOLD["focus_bt"][j], rtol=1e-5, atol=1e-5)
    s = onrm.normalize_sample(s)
    np.testing.assert_allclose(s.data["image_backtransform"], GOLD["norm_bt"][j], rtol=1e-5, atol=1e-4)
    np.testing.assert_array_equal(onrm.whiten_image(s.data["image"]), GOLD["net_in"][j])  # what the network is fed
    preds = Sample((S, S), {k: GOLD["pred_" + k][j] for k in PRED_CATS}, dict(PRED_CATS))
    preds.data["image_backtransform"] = s.data["image_backtransform"]
    preds = onrm.unnormalize_sample(preds)
    np.testing.assert_allclose(preds.data["image_backtransform"], GOLD["unnorm_bt"][j], rtol=1e-5, atol=1e-4)
    bt = preds.data.pop("image_backtransform")
    for k, cat in PRED_CATS.items():
        np.testing.assert_allclose(preds.data[k], GOLD["unnorm_" + k][j], rtol=1e-5, atol=1e-4, err_msg=k)
        got = apply_affine2d(bt, k, preds.data[k], cat)
        if k == "pose":
            assert min(np.abs(got - GOLD["final_pose"][j]).max(), np.abs(got + GOLD["final_pose"][j]).max()) < 2e-5
        else:
            np.testing.assert_allclose(got, GOLD["final_" + k][j], rtol=1e-4, atol=1e-3, err_msg=k)


@pytest.mark.parametrize("j", range(len(GOLD["cases"])))
def test_backtransform_through_flip_and_second_focus(j):
    c = cases.make_case(int(GOLD["cases"][j]))
    s = focus(c)
    do_flip, rot_dir = (int(x) for x in GOLD["flip_draws"][j])
    s = ogeo.horizontal_flip_and_rot_90(s, bool(do_flip), rot_dir)
    np.testing.assert_allclose(s.data["image_backtransform"], GOLD["flip_bt"][j], rtol=1e-5, atol=1e-4)
    s.categories = dict(CATS)
    s2 = focus(c, size=65, ext=1.3, insert=False, sample=s)
    np.testing.assert_allclose(s2.data["roi"], GOLD["again_roi"][j], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(s2.data["image_backtransform"], GOLD["again_bt"][j], rtol=1e-4, atol=1e-3)
