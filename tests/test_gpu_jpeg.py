"""GPU decode of JPEG source frames (b200aug_decode_jpeg_gray, nvJPEG) against the reference's decoder,
`cv2.imdecode(blob, 0)` (datasets/preprocessing.py:42-54).  Tolerance: 2 grey levels max abs (IDCT / colour-conversion
rounding differs between libjpeg-turbo and nvJPEG), mean abs < 0.3; then the decoded frames go through the fused
augmentation like any other device frames."""
import cv2
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def _blobs(sizes, color, quality=95, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i, (w, h) in enumerate(sizes):
        img = cases.make_image(rng, w, h, "smooth" if i % 3 else "noise")
        if color:
            img = np.stack([img, np.roll(img, 3, 1), 255 - img], -1)  # BGR
        ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality])
        assert ok
        out.append(buf.reshape(-1))
    return out


@pytest.mark.parametrize("color", [True, False])
def test_decode_matches_cv2(color):
    from trackertraincode_b200.datasets import preprocessing as pre

    sizes = [(450, 450)] * 6 + [(451, 333), (640, 480), (200, 180), (97, 131)]
    blobs = _blobs(sizes, color)
    frames = pre.imdecode_batch(blobs)
    assert len(frames) == len(blobs)
    for f, b, (w, h) in zip(frames, blobs, sizes):
        want = cv2.imdecode(b, 0)
        assert pre.jpeg_size(b)[:2] == (w, h) and tuple(f.shape) == (h, w) == want.shape and f.dtype == torch.uint8
        d = np.abs(f.cpu().numpy().astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 2 and d.mean() < 0.3, (w, h, d.max(), d.mean())
    stacked = pre.imdecode_batch(blobs[:6], stack=True)
    assert stacked.shape == (6, 450, 450) and all(torch.equal(stacked[i], frames[i]) for i in range(6))
    assert torch.equal(pre.imdecode(bytes(blobs[7])), frames[7])
    with pytest.raises(ValueError):
        pre.imdecode_batch(blobs, stack=True)
    with pytest.raises(Exception):
        pre.imdecode_batch([np.zeros(100, np.uint8)])  # not a JPEG


def test_decoded_frames_feed_the_fused_augmentation():
    from trackertraincode_b200.datasets import preprocessing as pre
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation

    n = 16
    blobs = _blobs([(450, 450)] * n, True, seed=5)
    frames = pre.imdecode_batch(blobs, stack=True)
    rng = np.random.default_rng(1)
    labs = [cases.make_labels(rng, 450, 450) for _ in range(n)]
    cats = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")
    data = {"image": frames}
    for k in ("roi", "coord", "pose", "pt3d_68"):
        data[k] = torch.from_numpy(np.stack([l[k] for l in labs])).cuda()
    b = Batch(Metadata((450, 450), n, "jpeg", None, {k: FieldCategory(v) for k, v in cats.items()}), data)
    aug = FusedPoseAugmentation(129, device="cuda", seed=1)
    torch.manual_seed(0)
    np.random.seed(0)
    d = aug.draw(n)
    out = aug(b, params=d)
    # same draws on the cv2-decoded frames: crops agree to the decode tolerance (area averaging only shrinks it)
    ref = torch.from_numpy(np.stack([cv2.imdecode(x, 0) for x in blobs])).cuda()
    out2 = aug(Batch(b.meta, {**{k: v for k, v in data.items() if k != "image"}, "image": ref}), params=d)
    assert out["image"].shape == (n, 1, 129, 129)
    noisy = d.photo.noise_apply.any(1) | d.photo.apply[:, 0] | d.photo.apply[:, 1]  # equalize / posterize amplify 1-LSB differences
    diff = (out["image"] - out2["image"]).abs().flatten(1).max(1).values.cpu()
    assert float(diff[~noisy].max()) <= 6.0 / 256, float(diff[~noisy].max())
