"""Channel-layout views, tensors/representation.py:15-28 (pure views, no kernel)."""
import torch


def ensure_image_nchw(img: torch.Tensor):
    assert not ((img.shape[-1] in (1, 3)) and (img.shape[-3] in (1, 3)))
    if img.shape[-3] in (1, 3):
        return img
    return img.swapaxes(-1, -2).swapaxes(-2, -3)


def ensure_image_nhwc(img: torch.Tensor):
    assert not ((img.shape[-1] in (1, 3)) and (img.shape[-3] in (1, 3)))
    if img.shape[-1] in (1, 3):
        return img
    return img.swapaxes(-3, -2).swapaxes(-2, -1)
