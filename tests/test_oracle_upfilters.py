"""The cubic / Lanczos up-filters (image_geometric_cv2.py:65-82, 105-119): the numpy models of OpenCV's fixed-point kernels
against the live cv2 binary, and the oracle's crop paths against outputs of the unmodified reference
(tests/golden/upfilter.npz, made by tests/golden/make_golden_upfilter.py).

cv2.resize(INTER_CUBIC) is the one place where the wheel's Intel IPP replaces OpenCV's kernel: the model is bit-exact
against OpenCV's own (IPP switched off) and within 1 grey level of IPP's on a few per cent of the pixels."""
import os

import cv2
import numpy as np
import pytest

import cases
import upfilter_cases as uc
from oracle import cv2_model, geometric as geo, normalization as nrm
from test_oracle_golden import host_cos_sin, to_sample

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FLAGS = {"cubic": cv2.INTER_CUBIC, "lanczos": cv2.INTER_LANCZOS4}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "upfilter.npz"))


@pytest.fixture()
def no_ipp():
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    yield
    cv2.ipp.setUseIPP(was)


def close_to_ipp(out, ref, what):
    d = np.abs(out.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.08, what


def test_taps():
    for kind, k in (("cubic", 4), ("lanczos", 8)):
        ofs, taps = cv2_model.resize_taps(57, 140, kind)
        assert taps.shape == (140, k) and np.abs(taps.sum(1) - 2048).max() <= 3  # (rounded taps: the sum is not normalised)
        tab = cv2_model.remap_table(kind)
        assert tab.shape == (32, 32, k, k) and (tab.sum((2, 3)) == 32768).all()
        # OpenCV's own quirk at phase (0, 0): the unit tap saturates to 32767 as a short and the fix-up hands the missing 1 to
        # the diagonal neighbour [k/2][k/2] (the model is bit-exact against cv2.warpAffine with exactly this table)
        o = k // 2 - 1
        assert tab[0, 0, o, o] == 32767 and tab[0, 0, o + 1, o + 1] == 1


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_resize_model_against_opencv_kernel(kind, no_ipp):
    rng = np.random.default_rng(1)
    for _ in range(8):
        h, w = (int(v) for v in rng.integers(4, 120, 2))
        dh, dw = (int(v) for v in rng.integers(max(h, w), 260, 2))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=FLAGS[kind])
        out = cv2_model.resize_cubic_or_lanczos_u8(img, dw, dh, kind)
        if kind == "lanczos":
            assert np.array_equal(out, ref), (h, w, dh, dw)
        else:  # (cv2's scalar tail columns and the odd near-tie: see the model's docstring)
            d = np.abs(out.astype(int) - ref.astype(int))
            assert d.max() <= 1 and (d > 0).sum() <= 3, (h, w, dh, dw)


def test_resize_model_against_the_wheel_as_shipped():
    """With IPP (if the wheel has it): Lanczos is OpenCV's kernel, cubic is IPP's -- within one grey level."""
    rng = np.random.default_rng(2)
    for _ in range(6):
        h, w = (int(v) for v in rng.integers(8, 120, 2))
        dh, dw = (int(v) for v in rng.integers(max(h, w), 260, 2))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(cv2_model.resize_cubic_or_lanczos_u8(img, dw, dh, "lanczos"), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LANCZOS4))
        close_to_ipp(cv2_model.resize_cubic_or_lanczos_u8(img, dw, dh, "cubic"), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC), (h, w, dh, dw))


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_warp_model_bit_exact(kind):
    rng = np.random.default_rng(3)
    for t in range(8):
        h, w = (int(v) for v in rng.integers(30, 200, 2))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ang, sc = rng.uniform(-0.7, 0.7), rng.uniform(0.6, 3.0)
        M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-40, 40)], [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-40, 40)]], np.float32)
        dw, dh = (int(v) for v in rng.integers(40, 200, 2))
        ref = cv2.warpAffine(img, M, (dw, dh), flags=FLAGS[kind], borderMode=cv2.BORDER_CONSTANT, borderValue=None)
        assert np.array_equal(cv2_model.warp_affine_cubic_or_lanczos_u8(img, M, dw, dh, kind), ref), t


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_focus_with_upfilter_matches_reference(gold, kind):
    for j, i in enumerate(uc.FOCUS_CASES):
        c = cases.make_case(i)
        s = nrm.offset_points_by_half_pixel(to_sample(c))
        params = geo.RoiFocusParams(c["scale"], c["angle"], c["translation"], host_cos_sin(c["angle"]))
        out, _ = geo.focus_roi(s, params, c["out_size"], use_model=True, upfilter=kind)
        ref = gold["focus_" + kind][j]
        if kind == "cubic" and i == 31 and bool(gold["ipp"]):  # the one case that goes through cv2.resize(INTER_CUBIC)
            close_to_ipp(out.data["image"][0], ref, f"case {i}")
        else:
            assert np.array_equal(out.data["image"][0], ref), f"case {i} ({kind})"


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_tensor_entries_with_upfilter_match_reference(gold, kind):
    for j, (i, entry, out_wh, g) in enumerate(uc.TENSOR_CASES):
        img = cases.make_case(i)["image"]
        if entry == "crop":
            out = geo.croprescale_image(img, g, out_wh, use_model=True, upfilter=kind)
        else:
            out = geo.affine_transform_image(img, gold["tensor_tr"][j], out_wh, use_model=True, upfilter=kind)
        ref = gold[f"tensor_{kind}_{j}"]
        assert out.shape == ref.shape
        if kind == "cubic" and entry == "crop" and bool(gold["ipp"]):
            close_to_ipp(out, ref, f"tensor case {j}")
        else:
            assert np.array_equal(out, ref), f"tensor case {j} ({kind})"
