"""oracle/photometric.py against hand-computed vectors of the specification in SURVEY.md 8(c).

kornia (the reference's dependency for this half, batch/intensity.py:9-64) is not installable here, so parity with
kornia itself is UNPINNED; these tests freeze the written-down algorithms so that the oracle -- and through it the CUDA
kernel -- cannot drift.
"""
import numpy as np

from oracle import photometric as P

F = np.float32


def test_posterize_hand_vector():
    x = np.array([[0, 15, 16, 200, 255]], F) / F(256)  # normalize_batch scale
    q = (x * F(255)).astype(np.uint8)                  # 0, 14, 15, 199, 254
    assert q.tolist() == [[0, 14, 15, 199, 254]]
    # the grey levels are exact; the final `/ 255.0` follows torch's GPU semantics (multiplication by float32(1/255))
    got = P.posterize(x, 4)
    assert np.array_equal(got, (np.array([[0, 0, 0, 192, 240]], F) * P.R255).astype(F))
    assert np.allclose(got, np.array([[0, 0, 0, 192, 240]], F) / F(255), rtol=0, atol=6e-8)
    got6 = P.posterize(x, 6)
    assert np.array_equal(got6, (np.array([[0, 12, 12, 196, 252]], F) * P.R255).astype(F))


def test_gamma_contrast_brightness_hand_vectors():
    x = np.array([[0.0, 0.25, 0.5, 1.0]], F)
    assert np.array_equal(P.gamma(x, 2.0), np.array([[0.0, 0.0625, 0.25, 1.0]], F))
    assert np.array_equal(P.gamma(x, 0.5), np.sqrt(x).astype(F))
    assert np.array_equal(P.contrast(x, 1.5), np.array([[0.0, 0.375, 0.75, 1.0]], F))  # clamp at 1
    assert np.array_equal(P.brightness(x, 0.75), np.array([[0.0, 0.0, 0.25, 0.75]], F))  # x + (b - 1), clamp at 0
    assert np.array_equal(P.brightness(x, 1.5), np.array([[0.5, 0.75, 1.0, 1.0]], F))


def test_equalize_hand_vector():
    # 4 grey levels, 4 pixels each: histogram h[10]=h[20]=h[30]=h[40]=4 on the x*255 scale
    im = np.repeat(np.array([10, 20, 30, 40], F), 4).reshape(4, 4)
    x = im / F(255)
    # step = (16 - 4) // 255 = 0 -> unchanged
    assert np.allclose(P.equalize(x), x, atol=1e-7)
    # 1024 pixels: 256 each -> step = (1024 - 256) // 255 = 3; lut[v] = (cumsum(h)[v-1] + 1) // 3 clamped to 255
    im = np.repeat(np.array([10, 20, 30, 40], F), 256).reshape(32, 32)
    got = P.equalize(im / F(255)) * F(255)
    want = {10: 0, 20: (256 + 1) // 3, 30: (512 + 1) // 3, 40: min((768 + 1) // 3, 255)}
    for v, w in want.items():
        assert np.allclose(got[im == v], w, atol=1e-4), (v, w, got[im == v][0])


def test_gaussian_kernel_and_blur():
    g = P.gaussian_kernel1d()
    t = np.arange(-2, 3, dtype=np.float64)
    ref = np.exp(-t * t / (2 * 1.5**2))
    ref /= ref.sum()
    assert np.allclose(g, ref, atol=1e-7) and abs(float(g.sum()) - 1) < 1e-6
    # constant image stays constant; an impulse spreads into the outer product of the kernel
    c = np.full((9, 11), 0.3, F)
    assert np.allclose(P.gaussian_blur(c), 0.3, atol=1e-6)
    imp = np.zeros((9, 9), F)
    imp[4, 4] = 1
    assert np.allclose(P.gaussian_blur(imp)[2:7, 2:7], np.outer(ref, ref), atol=1e-6)
    # reflect border (no edge repeat): column 0 sees x[2], x[1], x[0], x[1], x[2]
    row = np.tile(np.array([1, 2, 4, 8, 16, 32, 64], F), (7, 1))
    b = P.gaussian_blur(row)
    want0 = ref[0] * 4 + ref[1] * 2 + ref[2] * 1 + ref[3] * 2 + ref[4] * 4
    assert abs(float(b[3, 0]) - want0) < 1e-5


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    z = P.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    o = P.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in o] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = P.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(v) for v in p] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_noise_field_statistics_and_independence():
    a = P.noise_field(5, 0, 0, 129 * 129)
    b = P.noise_field(5, 1, 0, 129 * 129)
    c = P.noise_field(5, 0, 1, 129 * 129)
    assert abs(float(a.mean())) < 0.03 and abs(float(a.std()) - 1) < 0.03
    assert abs(float(np.corrcoef(a, b)[0, 1])) < 0.03 and abs(float(np.corrcoef(a, c)[0, 1])) < 0.03
    assert np.array_equal(a, P.noise_field(5, 0, 0, 129 * 129))  # counter based: reproducible


def test_stage_order_and_masks():
    rng = np.random.default_rng(0)
    x = rng.random((2, 1, 8, 8)).astype(F)
    p = P.PhotoParams(order=[P.OP_CONTRAST, P.OP_BRIGHTNESS], apply=np.zeros((2, 6), bool), bits=np.full(2, 5, np.int32),
                      gamma=np.ones(2, F), contrast=np.full(2, 1.5, F), brightness=np.full(2, 0.8, F),
                      noise_apply=np.zeros((2, 4), bool), seed=1)
    p.apply[0, P.OP_CONTRAST] = p.apply[0, P.OP_BRIGHTNESS] = True
    out = P.photometric_batch(x, p)
    assert np.array_equal(out[1], x[1])  # nothing selected for sample 1: unchanged (x already inside [0, 1])
    want = np.clip(np.clip(x[0] * F(1.5), 0, 1) + (F(0.8) - F(1)), 0, 1)
    assert np.array_equal(out[0], want)
    # noise without clip leaves [0, 1]; with OnlyClip it does not
    p.noise_apply[:, 3] = True
    p.clip = False
    assert P.photometric_batch(x, p).min() < 0
    p.clip = True
    o = P.photometric_batch(x, p)
    assert o.min() >= 0 and o.max() <= 1


def test_sampler_matches_pipeline_probabilities():
    rng = np.random.default_rng(3)
    n = 20000
    p = P.sample_photo_params(rng, n)
    assert len(p.order) == 4 and len(set(p.order)) == 4
    chosen = np.zeros(6, bool)
    chosen[p.order] = True
    freq = p.apply.mean(0)
    for k in range(6):
        assert abs(freq[k] - (P.DEFAULT_OP_PROB[k] if chosen[k] else 0)) < 0.01
    assert np.allclose(p.noise_apply.mean(0), P.DEFAULT_NOISE_PROB, atol=0.01)
    assert p.bits.min() >= 4 and p.bits.max() <= 5  # U(4,6) truncated
    assert 0.5 <= p.gamma.min() and p.gamma.max() <= 2.0
