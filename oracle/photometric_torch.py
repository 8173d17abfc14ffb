"""The photometric ops once more, written with exactly the torch primitives kornia calls -- TEST INFRASTRUCTURE.

kornia is not installed here (and cannot be: no network), so parity of the photometric half stays UNPINNED against
kornia itself.  What this file removes is the risk that `oracle/photometric.py` (numpy) mis-states what the torch
primitives underneath kornia do at bin edges / float32 roundings: every function below follows the source of
kornia 0.7.x as recalled (kornia/enhance/adjust.py: equalize -> _scale_channel / _build_lut, posterize, adjust_gamma,
adjust_contrast, adjust_brightness; kornia/filters/gaussian.py + filter.py: gaussian_blur2d -> filter2d_separable),
calling torch.histc, cumsum, div(rounding_mode="trunc"), gather, pow, clamp, F.pad(mode="reflect") + conv2d on float32
CPU tensors.  tests/test_oracle_photometric_torch.py asserts bitwise equality with oracle/photometric.py on random
inputs and on inputs sitting on bin edges (k/255, k/256, one ulp either side).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

KORNIA_VERSION_FOLLOWED = "0.7.x (kornia/enhance/adjust.py, kornia/filters/{gaussian,filter,kernels}.py), as recalled"


def _build_lut(histo: torch.Tensor, step: torch.Tensor) -> torch.Tensor:
    # kornia/enhance/adjust.py:_build_lut
    step_trunc = torch.div(step, 2, rounding_mode="trunc")
    lut = torch.div(torch.cumsum(histo, 0) + step_trunc, step, rounding_mode="trunc")
    lut = torch.cat([torch.zeros(1, device=lut.device, dtype=lut.dtype), lut[:-1]])
    return torch.clamp(lut, 0, 255)


def equalize(x: torch.Tensor) -> torch.Tensor:
    """kornia.enhance.equalize on one channel [H, W] in [0, 1] (adjust.py:_scale_channel)."""
    im = x * 255
    histo = torch.histc(im, bins=256, min=0, max=255)
    nonzero_histo = torch.reshape(histo[histo != 0], [-1])
    step = torch.div(torch.sum(nonzero_histo) - nonzero_histo[-1], 255, rounding_mode="trunc")
    if step == 0:
        result = im
    else:
        result = torch.gather(_build_lut(histo, step), 0, im.flatten().long()).reshape_as(im)
    return result / 255.0


def posterize(x: torch.Tensor, bits: int) -> torch.Tensor:
    """kornia.enhance.posterize (adjust.py): right shift then left shift of uint8(x * 255), through float divisions."""
    bits_t = torch.tensor(bits, device=x.device)
    if bits == 0:
        return torch.zeros_like(x)
    if bits == 8:
        return x.clone()
    shift = 8 - bits_t

    def _left_shift(inp, s):
        return ((inp * 255).to(torch.uint8) * (2**s)).to(inp.dtype) / 255.0

    def _right_shift(inp, s):
        return (inp * 255).to(torch.uint8) / (2**s).to(inp.dtype) / 255.0

    return _left_shift(_right_shift(x, shift), shift)


def gamma(x: torch.Tensor, g: float) -> torch.Tensor:
    """kornia.enhance.adjust_gamma(gain=1): clamp(gain * pow(x, gamma), 0, 1)."""
    return torch.clamp(1.0 * torch.pow(x, torch.tensor(g, dtype=torch.float32, device=x.device)), 0.0, 1.0)


def contrast(x: torch.Tensor, c: float) -> torch.Tensor:
    """kornia.enhance.adjust_contrast: clamp(x * factor, 0, 1)."""
    return torch.clamp(x * torch.tensor(c, dtype=torch.float32, device=x.device), 0.0, 1.0)


def brightness(x: torch.Tensor, b: float) -> torch.Tensor:
    """RandomBrightness -> adjust_brightness(x, b - 1): clamp(x + factor, 0, 1)."""
    return torch.clamp(x + (torch.tensor(b, dtype=torch.float32, device=x.device) - 1.0), 0.0, 1.0)


def gaussian_kernel1d(ksize: int = 5, sigma: float = 1.5, device="cpu") -> torch.Tensor:
    """kornia/filters/kernels.py:gaussian (odd window): exp(-x^2 / (2 sigma^2)) / sum, evaluated on `device`."""
    t = torch.arange(ksize, dtype=torch.float32, device=device) - ksize // 2
    g = torch.exp(-t.pow(2.0) / (2 * torch.tensor(sigma, dtype=torch.float32, device=device).pow(2.0)))
    return g / g.sum()


def gaussian_blur(x: torch.Tensor, ksize: int = 5, sigma: float = 1.5) -> torch.Tensor:
    """gaussian_blur2d(separable=True, border_type='reflect') on [H, W]: reflect pad + conv2d per axis (filter.py)."""
    k = gaussian_kernel1d(ksize, sigma, x.device)
    r = ksize // 2
    inp = x[None, None]
    out = F.conv2d(F.pad(inp, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, ksize))
    out = F.conv2d(F.pad(out, (0, 0, r, r), mode="reflect"), k.view(1, 1, ksize, 1))
    return out[0, 0]


def noise(x: torch.Tensor, z: torch.Tensor, std: float) -> torch.Tensor:
    """RandomGaussianNoise.apply_transform: input + noise * std (+ mean = 0)."""
    return x + z * std
