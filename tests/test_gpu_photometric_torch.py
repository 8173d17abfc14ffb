"""The photometric CUDA kernels against the torch primitives kornia runs in the reference's pipeline, evaluated ON THE GPU
(KorniaImageDistortions acts on CUDA tensors, pipelines.py:508-527): torch.histc / cumsum / div / gather (equalize), the
uint8 shifts (posterize), pow / clamp (gamma), mul / add + clamp, reflect pad + conv2d (gaussian blur) -- the
restatement in oracle/photometric_torch.py.  kornia itself is not installable here, so this is as close to the reference's
own arithmetic as the photometric half can be pinned."""
import numpy as np
import pytest
import torch

from oracle import photometric as opho, photometric_torch as T
from test_oracle_photometric_torch import IMAGES

pytestmark = pytest.mark.gpu
F32 = np.float32


def _run(images, order, **kw):
    """One stage-1 op (or none) through b200aug_photometric_f32 on a stack of equally sized images."""
    from trackertraincode_b200.datatransformation import _engine as E, batch as dtb

    n = images.shape[0]
    ones = torch.ones(n)
    p = E.PhotoParams(order, torch.ones(n, 6, dtype=torch.bool), kw.get("bits", torch.full((n,), 8, dtype=torch.int32)),
                      kw.get("gamma", ones), kw.get("contrast", ones), kw.get("brightness", ones), torch.zeros(n, 4, dtype=torch.bool),
                      clip=False)
    return dtb.photometric_f32(images, p)


def _stacks():
    by_shape = {}
    for x in IMAGES:
        by_shape.setdefault(x.shape, []).append(x)
    return [torch.from_numpy(np.stack(v)[:, None]).cuda() for v in by_shape.values()]


def test_gaussian_kernel_constants_are_torch_cuda():
    k = T.gaussian_kernel1d(device="cuda").cpu().numpy()
    assert np.array_equal(k, opho.gaussian_kernel1d())  # the constants compiled into the kernels


def test_equalize_and_posterize_bit_exact():
    for st in _stacks():
        got = _run(st, [opho.OP_EQUALIZE])
        for i in range(st.shape[0]):
            assert torch.equal(got[i, 0], T.equalize(st[i, 0])), "equalize"
        for bits in (4, 5, 6, 3, 7):
            got = _run(st, [opho.OP_POSTERIZE], bits=torch.full((st.shape[0],), bits, dtype=torch.int32))
            for i in range(st.shape[0]):
                assert torch.equal(got[i, 0], T.posterize(st[i, 0], bits)), f"posterize {bits}"


def test_contrast_brightness_bit_exact_gamma_two_ulp():
    rng = np.random.default_rng(4)
    for st in _stacks():
        n = st.shape[0]
        g, c, b = (torch.from_numpy(rng.uniform(lo, hi, n).astype(F32)) for lo, hi in ((0.5, 2.0), (0.7, 1.5), (0.7, 1.5)))
        got_c, got_b, got_g = _run(st, [opho.OP_CONTRAST], contrast=c), _run(st, [opho.OP_BRIGHTNESS], brightness=b), _run(st, [opho.OP_GAMMA], gamma=g)
        for i in range(n):
            assert torch.equal(got_c[i, 0], T.contrast(st[i, 0], float(c[i])))
            assert torch.equal(got_b[i, 0], T.brightness(st[i, 0], float(b[i])))
            a, w = got_g[i, 0].view(torch.int32).long(), T.gamma(st[i, 0], float(g[i])).view(torch.int32).long()
            assert int((a - w).abs().max()) <= 2, "gamma: powf of the kernel vs torch.pow on the GPU"


def test_gaussian_blur_within_rounding_of_conv2d():
    for st in _stacks():
        if min(st.shape[-2:]) < 5:
            continue
        got = _run(st, [opho.OP_BLUR])
        for i in range(st.shape[0]):
            want = T.gaussian_blur(st[i, 0])
            # cuDNN / native conv2d fuses and reorders the five multiply-adds per axis; values are <= 1
            assert float((got[i, 0] - want).abs().max()) <= 3e-7
