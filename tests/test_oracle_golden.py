"""Pin the numpy oracle against outputs of the unmodified reference (tests/golden/*.npz, made by
tests/golden/make_golden.py in the authoring container) and against the reference's own known-answer vectors."""
import os

import numpy as np
import pytest

import cases
from oracle import affine, geometric as geo, labels, normalization as nrm, pipeline
from oracle.geometric import Sample

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts", shapeparam="")
LABELS = ("roi", "coord", "pose", "pt3d_68", "shapeparam")


@pytest.fixture(scope="module")
def chain():
    return np.load(os.path.join(GOLD, "focus_chain.npz"))


@pytest.fixture(scope="module")
def algebra():
    return np.load(os.path.join(GOLD, "algebra.npz"))


# The 8 known-answer vectors of the reference: test/test_affine_img_trafo.py:49-69
VIEW_ROI_KAT = [
    ([-10, -10, 10, 10], 1.0, [-1.0, 0.0], 0.3, [-16, -10, 4, 10]),
    ([-10, -10, 10, 10], 1.0, [1.0, 0.0], 0.3, [-4, -10, 16, 10]),
    ([-10, -10, 10, 10], 1.0, [0.0, -1.0], 0.3, [-10, -16, 10, 4]),
    ([-10, -10, 10, 10], 1.0, [0.0, 1.0], 0.3, [-10, -4, 10, 16]),
    ([-10, -10, 10, 10], 2.0, [0.0, 0.0], 0.3, [-20, -20, 20, 20]),
    ([-10, -10, 10, 10], 2.0, [-1.0, 0.0], 0.3, [-36, -20, 4, 20]),
    ([-10, -10, 10, 10], 0.5, [0.0, 0.0], 0.3, [-5, -5, 5, 5]),
    ([-10, -10, 10, 10], 0.5, [-1.0, 0.0], 0.3, [-13, -5, -3, 5]),
]


@pytest.mark.parametrize("bbox,f,t,bbs,expected", VIEW_ROI_KAT)
def test_compute_view_roi_kat(bbox, f, t, bbs, expected):
    out = geo.compute_view_roi(np.float32(bbox), np.float32(f), np.float32(t), bbs)
    assert out.tolist() == expected


def to_sample(c) -> Sample:
    data = {"image": c["image"][..., None]}
    for k in LABELS:
        data[k] = c[k]
    return Sample(c["wh"], data, dict(CATS))


def rel_close(a, b, rtol=1e-5, atol=1e-5):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def quat_close(a, b, tol=1e-5, _unused=None):
    d = np.minimum(np.abs(a - b).max(-1), np.abs(a + b).max(-1))
    assert d.max() < tol


PRODUCTION_ANGLES = (0.0, float(np.float32(np.pi * 30.0 / 180.0)))


def host_cos_sin(angle):
    """What the reference's host code evaluates: torch.cos / torch.sin on a float32 CPU tensor (affine2d.py:46-47)."""
    import torch

    a = torch.tensor(float(angle), dtype=torch.float32)
    return np.float32(torch.cos(a).item()), np.float32(torch.sin(a).item())


@pytest.mark.parametrize("use_model,host_trig", [(False, True), (True, True), (False, False)])
def test_focus_chain_matches_reference(chain, use_model, host_trig):
    n_tr_bits = 0
    for i in range(cases.N_CASES):
        c = cases.make_case(i)
        if not host_trig and abs(float(c["angle"])) not in PRODUCTION_ANGLES:
            continue  # correctly rounded cos/sin differ from torch's by 1 ulp for ~5 % of arbitrary angles
        s = nrm.offset_points_by_half_pixel(to_sample(c))
        params = geo.RoiFocusParams(c["scale"], c["angle"], c["translation"], host_cos_sin(c["angle"]) if host_trig else None)
        s, inter = geo.focus_roi(s, params, c["out_size"], use_model=use_model)
        # integer work: bit exact
        assert np.array_equal(inter["view_roi"], chain["view_roi"][i]), f"case {i} view_roi"
        n_tr_bits += int(np.array_equal(inter["tr"], chain["tr"][i]))
        rel_close(inter["tr"], chain["tr"][i], 1e-6, 1e-5)
        # pixels: bit exact against the reference's cv2 output
        assert np.array_equal(s.data["image"][0], chain["focus_image"][i]), f"case {i} focus pixels"
        for k in LABELS:
            (quat_close if k == "pose" else rel_close)(s.data[k], chain["focus_" + k][i])
        s = geo.horizontal_flip_and_rot_90(s, c["do_flip"], c["rot_dir"])
        assert np.array_equal(s.data["image"][0], chain["flip_image"][i]), f"case {i} flip pixels"
        for k in LABELS:
            (quat_close if k == "pose" else rel_close)(s.data[k], chain["flip_" + k][i])
        s = nrm.normalize_sample(s)
        assert s.data["image"].dtype == np.float32
        assert np.array_equal(s.data["image"][0], chain["flip_image"][i].astype(np.float32) / 256)
        for k in LABELS:
            (quat_close if k == "pose" else rel_close)(s.data[k], chain["final_" + k][i], 1e-5, 1e-6)
    # the float32 transform itself is reproduced bit for bit (that is what keeps warpAffine's 1/32-px rounding aligned)
    assert n_tr_bits == (cases.N_CASES if host_trig else sum(abs(float(cases.make_case(i)["angle"])) in PRODUCTION_ANGLES for i in range(cases.N_CASES)))


def test_pipeline_wrapper_equals_stagewise(chain):
    for i in (0, 1, 4, 7, 31):
        c = cases.make_case(i)
        s, inter = pipeline.augment_sample(to_sample(c), c["scale"], c["angle"], c["translation"], c["do_flip"], c["rot_dir"], c["out_size"])
        assert np.array_equal(inter["view_roi"], chain["view_roi"][i])
        assert np.array_equal(s.data["image"][0], chain["flip_image"][i].astype(np.float32) / 256)
        rel_close(s.data["pt3d_68"], chain["final_pt3d_68"][i], 1e-5, 1e-6)


def test_affine_algebra(algebra):
    a = algebra
    n = len(a["angles"])
    for i in range(n):
        m = affine.trs(translations=a["translations"][i], angles=a["angles"][i], scales=a["scales_in"][i])
        if i % 3 == 0:
            m = affine.compose(m, affine.range_remap_2d([0.0, 0.0], [129, 129], [129, 0], [0, 129]))
        rel_close(m, a["mats"][i], 1e-6, 1e-5)
        b = affine.trs(translations=a["translations"][(i + 1) % n], angles=a["angles"][(i + 5) % n], scales=a["scales_in"][(i + 3) % n])
        rel_close(affine.compose(a["mats"][i], b), a["prods"][i], 1e-5, 1e-4)
        rel_close(affine.inv(a["mats"][i]), a["invs"][i], 1e-4, 1e-4)
        rel_close(affine.scales(a["mats"][i]), a["scales"][i], 1e-6, 0)
        rel_close(affine.det(a["mats"][i]), a["dets"][i], 1e-6, 1e-7)


def test_label_transforms(algebra):
    a = algebra
    assert np.array_equal(labels.FLIP_MAP, a["flip_map"])
    for i in range(len(a["angles"])):
        tr = a["mats"][i]
        rel_close(labels.transform_keypoints(tr, a["pts"][i]), a["pts_out"][i], 1e-5, 1e-4)
        rel_close(labels.transform_roi(tr, a["roi"][i]), a["roi_out"][i], 1e-5, 1e-4)
        rel_close(labels.transform_coord(tr, a["coord"][i]), a["coord_out"][i], 1e-5, 1e-4)
        quat_close(labels.transform_rot(tr, a["quat"][i]), a["quat_out"][i])
    # batched call == per-sample calls
    rel_close(labels.transform_keypoints(a["mats"], a["pts"]), a["pts_out"], 1e-5, 1e-4)
    quat_close(labels.transform_rot(a["mats"], a["quat"]), a["quat_out"])
    quat_close(labels.quat_mult(a["quat"], np.roll(a["quat"], 1, 0)), a["quat_mult"])
    rel_close(labels.quat_to_matrix(a["quat"]), a["quat_matrix"], 1e-5, 1e-6)
