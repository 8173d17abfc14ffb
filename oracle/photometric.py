"""Photometric chain (the kornia half of the reference) -- TEST INFRASTRUCTURE, **PARITY UNPINNED**.

The reference delegates these ops to kornia (`trackertraincode/datatransformation/batch/intensity.py:9-64`,
op list / probabilities / order in `trackertraincode/pipelines.py:508-532`).  kornia (unpinned,
`requirements.txt:7`) is not installed in any environment we can run and the reference has no test for this
half, so this file freezes the published algorithms as written down in SURVEY.md section 8(c); it is checked
against hand-computed vectors in tests/test_oracle_photometric.py, not against kornia itself.

All ops act on one float32 image [H, W] with values in [0, 1] and take their sampled parameters explicitly.
Gaussian noise is a counter-based Philox4x32-10 stream keyed by (seed; sample id, stage, pixel group) so the CUDA
kernel and this file draw identical numbers.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

F32 = np.float32

OP_EQUALIZE, OP_POSTERIZE, OP_GAMMA, OP_CONTRAST, OP_BRIGHTNESS, OP_BLUR = range(6)
NUM_OPS = 6
NUM_NOISE = 4
# pipelines.py:510-527
DEFAULT_OP_PROB = (0.2, 0.01, 0.2, 0.2, 0.2, 0.1)
DEFAULT_RANDOM_APPLY = 4
DEFAULT_NOISE_STD = (4.0 / 255.0, 16.0 / 255.0, 32.0 / 255.0, 64.0 / 255.0)
DEFAULT_NOISE_PROB = (0.25, 0.25**2, 0.25**3, 0.25**4)
POSTERIZE_BITS_RANGE = (4.0, 6.0)
GAMMA_RANGE = (0.5, 2.0)
CONTRAST_RANGE = (0.7, 1.5)
BRIGHTNESS_RANGE = (0.7, 1.5)
BLUR_KSIZE, BLUR_SIGMA = 5, 1.5


# kornia's `x / 255.0`: the reference runs these ops on CUDA tensors (pipelines.py:508-527), where torch evaluates a tensor
# divided by a host scalar as a multiplication with the float32 reciprocal -- one ulp away from the true quotient now and
# then.  tests/test_gpu_photometric_torch.py checks the CUDA kernels bit for bit against the torch primitives on the GPU.
R255 = F32(1.0) / F32(255.0)

# ----------------------------------------------------------------------------- point ops


def equalize(x: np.ndarray) -> np.ndarray:
    """kornia.enhance.equalize on one channel: 256-bin histogram of x*255 over [0, 255], TF-style LUT."""
    x = np.asarray(x, F32)
    im = x * F32(255.0)
    pos = (im / F32(255.0)) * F32(256.0)  # torch.histc bin position, float32
    idx = pos.astype(np.int64)
    idx = np.where(idx == 256, 255, idx)  # the max value belongs to the last bin
    valid = (im >= 0) & (im <= 255)
    h = np.bincount(idx[valid].ravel(), minlength=256)[:256].astype(np.int64)
    nz = h[h != 0]
    step = (int(nz.sum()) - int(nz[-1])) // 255 if nz.size else 0
    if step == 0:
        return (im * R255).astype(F32)
    lut = (np.cumsum(h) + step // 2) // step
    lut = np.clip(np.concatenate([[0], lut[:-1]]), 0, 255).astype(F32)
    return (lut[np.clip(im.astype(np.int64), 0, 255)] * R255).astype(F32)


def posterize(x: np.ndarray, bits: int) -> np.ndarray:
    """kornia.enhance.posterize, literally: _left_shift(_right_shift(x, s), s) with s = 8 - bits, where
    _right_shift = uint8(x * 255) / 2**s / 255 and _left_shift = uint8(x * 255) * 2**s / 255 (float32, `/ 255` as R255)."""
    shift = 8 - int(bits)
    q = (np.asarray(x, F32) * F32(255.0)).astype(np.uint8)
    right = ((q.astype(F32) / F32(2**shift)) * R255).astype(F32)
    left = ((right * F32(255.0)).astype(np.uint8).astype(np.int64) * (2**shift)) & 255
    return (left.astype(F32) * R255).astype(F32)


def gamma(x: np.ndarray, g: float) -> np.ndarray:
    """kornia.enhance.adjust_gamma(gain=1): clamp(x**g, 0, 1)."""
    return np.clip(np.power(np.asarray(x, F32), F32(g)), 0, 1).astype(F32)


def contrast(x: np.ndarray, c: float) -> np.ndarray:
    """kornia RandomContrast -> adjust_contrast: clamp(x*c, 0, 1)."""
    return np.clip(np.asarray(x, F32) * F32(c), 0, 1).astype(F32)


def brightness(x: np.ndarray, b: float) -> np.ndarray:
    """kornia RandomBrightness -> adjust_brightness(x, b-1): clamp(x + (b-1), 0, 1)."""
    return np.clip(np.asarray(x, F32) + (F32(b) - F32(1.0)), 0, 1).astype(F32)


def gaussian_kernel1d(ksize=BLUR_KSIZE, sigma=BLUR_SIGMA) -> np.ndarray:
    t = np.arange(ksize, dtype=F32) - F32(ksize // 2)
    g = np.exp(-(t * t) / F32(2.0 * sigma * sigma)).astype(F32)
    return (g / g.sum(dtype=F32)).astype(F32)


def _reflect(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def gaussian_blur(x: np.ndarray, ksize=BLUR_KSIZE, sigma=BLUR_SIGMA) -> np.ndarray:
    """kornia gaussian_blur2d(separable, border 'reflect'): horizontal pass then vertical pass, float32."""
    x = np.asarray(x, F32)
    h, w = x.shape
    g = gaussian_kernel1d(ksize, sigma)
    r = ksize // 2
    xs = np.arange(w)
    ys = np.arange(h)
    tmp = np.zeros_like(x)
    for i in range(ksize):
        term = g[i] * x[:, _reflect(xs + i - r, w)]
        tmp = term if i == 0 else tmp + term
    out = np.zeros_like(x)
    for j in range(ksize):
        term = g[j] * tmp[_reflect(ys + j - r, h), :]
        out = term if j == 0 else out + term
    return out.astype(F32)


# ----------------------------------------------------------------------------- Philox noise

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
NOISE_DOMAIN = 0x6E6F6973  # "nois": 4th counter word, xor the high 32 bits of the sample id


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Standard Philox4x32-10; counters are uint32 arrays (broadcast), key two python ints. Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & _MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & _MASK, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & _MASK, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _box_muller(xa, xb):
    u1 = ((xa >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0**-24  # (0, 1]
    u2 = (xb >> np.uint32(8)).astype(np.float64) * 2.0**-24  # [0, 1)
    r = np.sqrt(-2.0 * np.log(u1))
    return (r * np.cos(2.0 * np.pi * u2)).astype(F32), (r * np.sin(2.0 * np.pi * u2)).astype(F32)


def noise_field(seed: int, sample_id: int, stage: int, npix: int) -> np.ndarray:
    """Standard-normal field for one image and noise stage, flat [npix].

    One Philox call (counter = (g, stage, sample_id, NOISE_DOMAIN), key = seed) yields 4 normals that go to pixels
    g, g+Q, g+2Q, g+3Q with Q = ceil(npix/4) -- the stride layout lets a warp store 32 consecutive floats."""
    q = (npix + 3) // 4
    g = np.arange(q, dtype=np.uint32)
    x0, x1, x2, x3 = philox4x32_10(g, np.uint32(stage), np.uint32(sample_id & 0xFFFFFFFF), np.uint32(NOISE_DOMAIN ^ ((sample_id >> 32) & 0xFFFFFFFF)),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    z0, z1 = _box_muller(x0, x1)
    z2, z3 = _box_muller(x2, x3)
    return np.concatenate([z0, z1, z2, z3])[:npix]


# ----------------------------------------------------------------------------- the chain


@dataclass
class PhotoParams:
    """Everything random about one call of the two KorniaImageDistortions stages, for a batch of B images."""

    order: Sequence[int]  # stage-1 op ids in application order (random_apply draws 4 of 6 per call)
    apply: np.ndarray  # bool [B, NUM_OPS], indexed by op id
    bits: np.ndarray  # int32 [B]
    gamma: np.ndarray  # float32 [B]
    contrast: np.ndarray  # float32 [B]
    brightness: np.ndarray  # float32 [B]
    noise_apply: np.ndarray  # bool [B, NUM_NOISE]
    noise_std: Sequence[float] = DEFAULT_NOISE_STD
    seed: int = 0
    sample_offset: int = 0  # id of the batch's first sample in the noise stream
    clip: bool = True  # OnlyClip(p=1)
    noise_clip: Sequence[bool] = (False,) * NUM_NOISE  # RandomGaussianNoiseWithClipping (intensity.py:43-53) per stage

    def slice(self, lo, hi):
        return PhotoParams(self.order, self.apply[lo:hi], self.bits[lo:hi], self.gamma[lo:hi], self.contrast[lo:hi],
                           self.brightness[lo:hi], self.noise_apply[lo:hi], self.noise_std, self.seed,
                           self.sample_offset + lo, self.clip, self.noise_clip)


def sample_photo_params(rng: np.random.Generator, B: int, seed: int = 0, sample_offset: int = 0,
                        op_prob=DEFAULT_OP_PROB, random_apply=DEFAULT_RANDOM_APPLY,
                        noise_prob=DEFAULT_NOISE_PROB, noise_std=DEFAULT_NOISE_STD) -> PhotoParams:
    """The sampling kornia does per call (pipelines.py:510-527): choose `random_apply` of the 6 ops (one draw per
    call, shared by the batch), per-op per-sample Bernoulli masks, uniform factors."""
    order = rng.permutation(NUM_OPS)[:random_apply].tolist()
    apply = rng.random((B, NUM_OPS)) < np.asarray(op_prob)[None, :]
    chosen = np.zeros(NUM_OPS, bool)
    chosen[order] = True
    apply &= chosen[None, :]
    return PhotoParams(
        order=order,
        apply=apply,
        bits=rng.uniform(*POSTERIZE_BITS_RANGE, B).astype(np.int32),  # kornia truncates the sampled float
        gamma=rng.uniform(*GAMMA_RANGE, B).astype(F32),
        contrast=rng.uniform(*CONTRAST_RANGE, B).astype(F32),
        brightness=rng.uniform(*BRIGHTNESS_RANGE, B).astype(F32),
        noise_apply=rng.random((B, NUM_NOISE)) < np.asarray(noise_prob)[None, :],
        noise_std=tuple(noise_std),
        seed=seed,
        sample_offset=sample_offset,
    )


def apply_stage1(x: np.ndarray, p: PhotoParams, b: int) -> np.ndarray:
    for op in p.order:
        if not p.apply[b, op]:
            continue
        if op == OP_EQUALIZE:
            x = equalize(x)
        elif op == OP_POSTERIZE:
            x = posterize(x, int(p.bits[b]))
        elif op == OP_GAMMA:
            x = gamma(x, p.gamma[b])
        elif op == OP_CONTRAST:
            x = contrast(x, p.contrast[b])
        elif op == OP_BRIGHTNESS:
            x = brightness(x, p.brightness[b])
        elif op == OP_BLUR:
            x = gaussian_blur(x)
    return x


def apply_stage2(x: np.ndarray, p: PhotoParams, b: int) -> np.ndarray:
    """4 x RandomGaussianNoise (no clip in between) then OnlyClip (pipelines.py:521-527, intensity.py:56-64)."""
    h, w = x.shape
    for s in range(NUM_NOISE):
        if p.noise_apply[b, s]:
            z = noise_field(p.seed, p.sample_offset + b, s, h * w).reshape(h, w)
            x = (x + F32(p.noise_std[s]) * z).astype(F32)
            if p.noise_clip[s]:  # the clipping variant clamps the samples it was applied to (intensity.py:52)
                x = np.clip(x, 0, 1).astype(F32)
    if p.clip:
        x = np.clip(x, 0, 1).astype(F32)
    return x


def photometric_batch(images: np.ndarray, p: PhotoParams) -> np.ndarray:
    """images float32 [B, 1, H, W] in [0, 1] -> same shape; both KorniaImageDistortions stages."""
    out = np.empty_like(images, dtype=F32)
    for b in range(images.shape[0]):
        out[b, 0] = apply_stage2(apply_stage1(np.asarray(images[b, 0], F32), p, b), p, b)
    return out
