"""oracle/cv2_model.py (the specification the CUDA kernels implement) against the live OpenCV binary, bit for bit.

OpenCV is the third-party dependency behind the reference's pixel path (image_geometric_cv2.py:65-135); the same wheel
(opencv-python-headless 4.13.0) is installed here and on the GPU box, so these run everywhere without a GPU.
"""
import cv2
import numpy as np
import pytest

from oracle import cv2_model as M


def images(rng, h, w):
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    smooth = np.clip((np.sin(x / 17.0) + np.cos(y / 23.0) + 2.0) / 4.0 * 255.0 + rng.normal(0, 8, (h, w)), 0, 255)
    return [rng.integers(0, 256, (h, w), dtype=np.uint8), smooth.astype(np.uint8)]


@pytest.mark.parametrize("seed", range(6))
def test_warp_affine_linear_bit_exact(seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(60, 200)), int(rng.integers(60, 200))
    for img in images(rng, h, w):
        ang = rng.uniform(-np.pi, np.pi)
        s = rng.uniform(0.5, 2.0)
        Mx = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-30, 30)],
                       [s * np.sin(ang), s * np.cos(ang), rng.uniform(-30, 30)]], np.float32).astype(np.float64)
        dw, dh = int(rng.integers(20, 220)), int(rng.integers(20, 220))
        want = cv2.warpAffine(img, Mx, (dw, dh), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        got = M.warp_affine_linear_u8(img, Mx, dw, dh)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(6))
def test_resize_area_general_bit_exact(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(4):
        sh, sw = int(rng.integers(130, 420)), int(rng.integers(130, 420))
        dh, dw = int(rng.integers(40, 130)), int(rng.integers(40, 130))
        for img in images(rng, sh, sw):
            want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
            assert np.array_equal(M.resize_area_u8(img, dw, dh), want)


@pytest.mark.parametrize("fx,fy", [(2, 2), (3, 3), (2, 3), (4, 2), (1, 2)])
def test_resize_area_integer_factor_bit_exact(fx, fy):
    rng = np.random.default_rng(7)
    dw, dh = 43, 37
    for img in images(rng, dh * fy, dw * fx):
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(M.resize_area_u8(img, dw, dh), want)


@pytest.mark.parametrize("seed", range(4))
def test_resize_linear_upscale_bit_exact(seed):
    rng = np.random.default_rng(200 + seed)
    sh, sw = int(rng.integers(20, 128)), int(rng.integers(20, 128))
    dh, dw = int(rng.integers(129, 260)), int(rng.integers(129, 260))
    for img in images(rng, sh, sw):
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(M.resize_linear_u8(img, dw, dh), want)


def test_identity_sizes_are_copies():
    img = np.arange(35, dtype=np.uint8).reshape(5, 7)
    assert np.array_equal(M.resize_area_u8(img, 7, 5), img)
    assert np.array_equal(M.resize_linear_u8(img, 7, 5), img)
