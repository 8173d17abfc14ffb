"""oracle/photometric.py (numpy) against oracle/photometric_torch.py (the torch primitives kornia itself calls:
torch.histc, cumsum, div(trunc), gather, pow, clamp, reflect pad + conv2d) -- bitwise, on random images and on images
whose values sit on the bin edges of the histogram / the uint8 quantiser (k/255, k/256, one ulp either side)."""
import numpy as np
import pytest
import torch

from oracle import photometric as P, photometric_torch as T

F32 = np.float32


def edge_images():
    k = np.arange(256, dtype=F32)
    base = [k / F32(255.0), k / F32(256.0), (k + F32(0.5)) / F32(255.0), k * F32(1.0 / 255.0), k * F32(1.0 / 256.0),
            (k * F32(255.0 / 256.0)) / F32(255.0)]   # histogram bin edges in the x*255 domain are k * 255/256
    out = []
    for b in base:
        for d in (0, 1, -1):
            v = b.copy()
            if d:
                v = np.nextafter(v, F32(2.0) if d > 0 else F32(-1.0)).astype(F32)
            out.append(np.clip(v, 0, 1).astype(F32).reshape(16, 16))
    return out


def random_images():
    rng = np.random.default_rng(0)
    imgs = [rng.random((37, 53)).astype(F32), (rng.integers(0, 256, (64, 48)) / F32(256.0)).astype(F32),
            (rng.integers(0, 256, (33, 65)) / F32(255.0)).astype(F32), np.full((9, 9), 0.25, F32),
            (np.round(rng.random((40, 40)) * 5) / 5).astype(F32)]
    imgs.append(np.clip(rng.normal(0.5, 0.2, (129, 129)), 0, 1).astype(F32))
    return imgs


IMAGES = edge_images() + random_images()


def _ulps(a, b):
    return int(np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max())


@pytest.mark.parametrize("i", range(len(IMAGES)))
def test_equalize_same_bins_same_lut(i):
    """Histogram bins, step and LUT entries are integers: identical.  The final `/ 255.0` is a true division on the CPU and a
    multiplication by float32(1/255) on the GPU (which the oracle follows): the float results are within one ulp."""
    x = IMAGES[i]
    a, w = P.equalize(x), T.equalize(torch.from_numpy(x)).numpy()
    assert _ulps(a, w) <= 1


@pytest.mark.parametrize("bits", [1, 3, 4, 5, 6, 7])
def test_posterize_same_levels(bits):
    for x in IMAGES:
        a, w = P.posterize(x, bits), T.posterize(torch.from_numpy(x), bits).numpy()
        assert _ulps(a, w) <= 1, bits


def test_point_ops_bitwise():
    rng = np.random.default_rng(1)
    for x in IMAGES:
        for _ in range(3):
            g, c, b = (F32(rng.uniform(0.5, 2.0)), F32(rng.uniform(0.7, 1.5)), F32(rng.uniform(0.7, 1.5)))
            t = torch.from_numpy(x)
            assert np.array_equal(P.contrast(x, c), T.contrast(t, float(c)).numpy())
            assert np.array_equal(P.brightness(x, b), T.brightness(t, float(b)).numpy())
            # pow: numpy and torch may call different libm / SIMD routines: at most one ulp apart
            a, w = P.gamma(x, g), T.gamma(t, float(g)).numpy()
            assert np.abs(a.view(np.int32).astype(np.int64) - w.view(np.int32).astype(np.int64)).max() <= 1


def test_gaussian_kernel_and_blur():
    # torch's CPU exp and numpy's differ by one ulp at exp(-8/9); torch's CUDA exp -- what the reference's pipeline runs --
    # equals numpy's (tests/test_gpu_photometric_torch.py asserts that bit for bit)
    a, w = P.gaussian_kernel1d(), T.gaussian_kernel1d().numpy()
    assert np.abs(a.view(np.int32).astype(np.int64) - w.view(np.int32).astype(np.int64)).max() <= 2
    for x in IMAGES:
        if min(x.shape) < 3:
            continue
        a, w = P.gaussian_blur(x), T.gaussian_blur(torch.from_numpy(x)).numpy()
        # conv2d may fuse multiply-adds / reorder the five taps: a few ulp of a value <= 1, never more
        assert np.abs(a - w).max() <= 2.4e-7, np.abs(a - w).max()


def test_noise_formula():
    rng = np.random.default_rng(3)
    x, z = rng.random((20, 20)).astype(F32), rng.standard_normal((20, 20)).astype(F32)
    for std in P.DEFAULT_NOISE_STD:
        want = T.noise(torch.from_numpy(x), torch.from_numpy(z), std).numpy()
        assert np.array_equal((x + F32(std) * z).astype(F32), want)
