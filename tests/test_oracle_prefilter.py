"""The anti-alias down-filters (image_geometric_cv2.py:47-82): the numpy models of cv2.GaussianBlur / cv2.sepFilter2D against
the live cv2 binary, and the oracle's filtered crop paths against outputs of the unmodified reference
(tests/golden/prefilter.npz, made by tests/golden/make_golden_prefilter.py)."""
import os

import cv2
import numpy as np
import pytest

import cases
import prefilter_cases as pc
from oracle import cv2_model, geometric as geo, normalization as nrm
from test_oracle_golden import host_cos_sin, to_sample

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "prefilter.npz"))


def test_hamming_window_is_scipys():
    import scipy.signal.windows

    for n in range(3, 65, 2):
        sf = 0.99 if n == 3 else 2.0 / (n - 1) * 0.999
        ours = cv2_model.hamming_kernel(sf)
        ref = scipy.signal.windows.hamming(n)
        ref /= np.sum(ref)
        assert len(ours) == n and np.array_equal(ours, ref), n  # bit for bit: cv2's symmetry test looks at the last bit


def test_gaussian_taps_sum_to_one_and_mirror():
    for sigma in np.linspace(0.5001, 10.0, 97):
        k = cv2_model.gaussian_kernel_fixed(float(sigma))
        assert k.sum() == 256 and np.array_equal(k, k[::-1]) and len(k) == cv2_model.gaussian_ksize_u8(float(sigma))


@pytest.mark.parametrize("seed", range(6))
def test_gaussian_blur_model_bit_exact(seed):
    rng = np.random.default_rng(seed)
    for _ in range(8):
        h, w = rng.integers(3, 260, 2)
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if seed % 2:
            img = ((rng.random((h, w)) < 0.5) * 255).astype(np.uint8)
        sigma = 0.5 / rng.uniform(0.05, 0.999)
        ref = cv2.GaussianBlur(img, (0, 0), sigma, sigma, borderType=cv2.BORDER_REPLICATE)
        assert np.array_equal(cv2_model.gaussian_blur_u8(img, sigma, sigma, "replicate"), ref)
        # the call as the reference writes it (image_geometric_cv2.py:50): sigmaY = 1, reflect-101 border
        ref = cv2.GaussianBlur(img, (0, 0), sigma, sigma, cv2.BORDER_REPLICATE)
        assert np.array_equal(cv2_model.antialias_prefilter_u8(img, 0.5 / sigma, "gaussian"), ref)


@pytest.mark.parametrize("seed", range(6))
def test_hamming_filter_model(seed):
    """Bit-exact on the columns cv2 filters with its vector loops; its scalar tail loops (the last width mod 32 columns
    at most, depending on the CPU's vector width) round differently: there, at most 1 LSB on a few pixels."""
    rng = np.random.default_rng(100 + seed)
    for _ in range(6):
        h, w = rng.integers(70, 300, 2)
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        k = cv2_model.hamming_kernel(rng.uniform(0.07, 0.999))
        ref = cv2.sepFilter2D(img, -1, k, k)
        out = cv2_model.sep_filter_u8(img, k)
        d = np.abs(out.astype(int) - ref.astype(int))
        body = w - w % 64
        assert d[:, :body].max() == 0
        assert d.max() <= 1 and (d > 0).mean() < 1e-3


@pytest.mark.parametrize("kind", pc.FILTERS)
def test_focus_with_prefilter_matches_reference(gold, kind):
    for j, i in enumerate(pc.FOCUS_CASES):
        c = cases.make_case(i)
        s = nrm.offset_points_by_half_pixel(to_sample(c))
        params = geo.RoiFocusParams(c["scale"], c["angle"], c["translation"], host_cos_sin(c["angle"]))
        for use_model in (False, True):
            out, _ = geo.focus_roi(s, params, c["out_size"], use_model=use_model, downfilter=kind)
            d = np.abs(out.data["image"][0].astype(int) - gold["focus_" + kind][j].astype(int))
            if kind == "gaussian" or not use_model:
                assert d.max() == 0, f"case {i} ({kind}, model={use_model})"
            else:  # (cv2's scalar tail columns, see test_hamming_filter_model)
                assert d.max() <= 1 and (d > 0).mean() < 2e-3, f"case {i} ({kind})"


@pytest.mark.parametrize("kind", pc.FILTERS)
def test_tensor_entries_with_prefilter_match_reference(gold, kind):
    for j, (i, entry, out_wh, g) in enumerate(pc.TENSOR_CASES):
        img = cases.make_case(i)["image"]
        if entry == "crop":
            out = geo.croprescale_image(img, g, out_wh, use_model=True, downfilter=kind)
        else:
            out = geo.affine_transform_image(img, gold["tensor_tr"][j], out_wh, use_model=True, downfilter=kind)
        ref = gold[f"tensor_{kind}_{j}"]
        assert out.shape == ref.shape
        d = np.abs(out.astype(int) - ref.astype(int))
        if kind == "gaussian":
            assert d.max() == 0, f"tensor case {j}"
        else:
            assert d.max() <= 1 and (d > 0).mean() < 2e-3, f"tensor case {j}"
