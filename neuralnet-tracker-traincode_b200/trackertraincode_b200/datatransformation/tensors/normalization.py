"""whiten / unwhiten, tensors/normalization.py:19-24.  Inside the training path the subtraction is fused into the crop
kernel (B200AUG_F_WHITEN); these are the stand-alone forms eval.py:199,230 and scripts/show_train_test_splits.py:27 call."""
import torch


def whiten_image(image: torch.Tensor):
    return image.sub(0.5)


def unwhiten_image(image: torch.Tensor):
    return image.add(0.5)
