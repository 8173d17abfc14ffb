"""The anti-alias down-filters 'gaussian' / 'hamming' (image_geometric_cv2.py:47-82) on the GPU, through the mirror API,
against outputs of the unmodified reference (tests/golden/prefilter.npz) and against the oracle's models.

gaussian: exact integer arithmetic -> bit-exact against the reference.  hamming: cv2 filters in float32 with vector loops
and (for the last few columns of an image) scalar tail loops that round differently; kernel and oracle model follow the
vector loops -> bit-exact against each other, within 1 LSB on < 0.2 % of the pixels against the reference."""
import os

import numpy as np
import pytest
import torch

import cases
import prefilter_cases as pc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "prefilter.npz"))


def focus_batch(idx):
    """Cases of one frame size as a stacked batch for the mirror's RandomFocusRoi (behind offset_points_by_half_pixel)."""
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

    cs = [cases.make_case(i) for i in idx]
    w, h = cs[0]["wh"]
    cats = {"image": FieldCategory.image, "roi": FieldCategory.roi, "pt3d_68": FieldCategory.points}
    data = {"image": torch.from_numpy(np.stack([c["image"] for c in cs])[:, None]).cuda(),
            "roi": torch.from_numpy(np.stack([c["roi"] for c in cs])).cuda(),
            "pt3d_68": torch.from_numpy(np.stack([c["pt3d_68"] for c in cs])).cuda()}
    return cs, Batch(Metadata((w, h), len(cs), None, None, cats), data)


def check(kind, out, ref, model, what):
    assert out.shape == ref.shape, what
    assert np.array_equal(out, model), f"{what}: kernel vs oracle model"
    d = np.abs(out.astype(int) - ref.astype(int))
    if kind == "gaussian":
        assert d.max() == 0, f"{what}: kernel vs reference"
    else:
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, f"{what}: kernel vs reference"


@pytest.mark.parametrize("kind", pc.FILTERS)
def test_focus_with_prefilter(gold, kind):
    from oracle import geometric as geo, normalization as nrm
    from test_oracle_golden import host_cos_sin, to_sample
    from trackertraincode_b200.datatransformation import batch as dtb

    by_size = {}
    for j, i in enumerate(pc.FOCUS_CASES):
        by_size.setdefault(cases.make_case(i)["wh"], []).append((j, i))
    for wh, items in by_size.items():
        cs, b = focus_batch([i for _, i in items])
        b = dtb.offset_points_by_half_pixel(b)
        params = dtb.RoiFocusRandomizationParameters(
            scales=torch.tensor([float(c["scale"]) for c in cs]), angles=torch.tensor([float(c["angle"]) for c in cs]),
            translations=torch.from_numpy(np.stack([c["translation"] for c in cs])), upfilter="linear", downfilter=kind)
        aug = dtb.RandomFocusRoi(cases.OUT_SIZE)
        aug.make_randomization_parameters = lambda B, params=params: params
        res = aug(b)
        aug.status.flush()
        img = res["image"].cpu().numpy()[:, 0]
        for n, (j, i) in enumerate(items):
            c = cs[n]
            s = nrm.offset_points_by_half_pixel(to_sample(c))
            want, _ = geo.focus_roi(s, geo.RoiFocusParams(c["scale"], c["angle"], c["translation"], host_cos_sin(c["angle"])),
                                    c["out_size"], use_model=True, downfilter=kind)
            check(kind, img[n], gold["focus_" + kind][j], want.data["image"][0], f"case {i}")


@pytest.mark.parametrize("kind", pc.FILTERS)
def test_tensor_entries_with_prefilter(gold, kind):
    from oracle import geometric as geo
    from trackertraincode_b200.datatransformation import tensors as dtt
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    for j, (i, entry, out_wh, g) in enumerate(pc.TENSOR_CASES):
        frame = cases.make_case(i)["image"]
        img = torch.from_numpy(frame[None].copy()).cuda()
        if entry == "crop":
            out = dtt.croprescale_image_cv2(img, torch.tensor(g, dtype=torch.int32), out_wh, downfilter=kind)
            model = geo.croprescale_image(frame, g, out_wh, use_model=True, downfilter=kind)
        else:
            tr = gold["tensor_tr"][j]
            out = dtt.affine_transform_image_cv2(img, Affine2d(torch.from_numpy(tr.copy())), out_wh, downfilter=kind)
            model = geo.affine_transform_image(frame, tr, out_wh, use_model=True, downfilter=kind)
        check(kind, out.cpu().numpy()[0], gold[f"tensor_{kind}_{j}"], model, f"tensor case {j}")


def test_prefilter_too_wide_is_reported():
    """A 640 x 480 frame shrunk to 12 x 12 asks for a 161-tap Gaussian: beyond B200AUG_PREFILTER_MAX_TAPS -> NativeError,
    not silent zeros."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import tensors as dtt

    img = torch.from_numpy(cases.make_case(4)["image"][None].copy()).cuda()
    with pytest.raises(N.NativeError):
        dtt.croprescale_image_cv2(img, torch.tensor([0, 0, 640, 480], dtype=torch.int32), (12, 12), downfilter="gaussian")
    with pytest.raises(NotImplementedError):
        dtt.croprescale_image_cv2(img, torch.tensor([0, 0, 640, 480], dtype=torch.int32), (64, 64), downfilter="box")


def test_random_crops_against_the_models():
    """Random boxes / sizes / angles beyond the golden set, kernel vs oracle models (bit-exact for both filters)."""
    from oracle import geometric as geo
    from trackertraincode_b200.datatransformation import tensors as dtt
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    rng = np.random.default_rng(5)
    for t in range(24):
        h, w = (int(v) for v in rng.integers(120, 420, 2))
        frame = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img = torch.from_numpy(frame[None].copy()).cuda()
        ow, oh = (int(v) for v in rng.integers(24, 100, 2))
        kind = pc.FILTERS[t % 2]
        if t % 3:
            x0, y0 = (int(v) for v in rng.integers(-30, 40, 2))
            x1, y1 = x0 + int(rng.integers(ow + 8, w + 40)), y0 + int(rng.integers(oh + 8, h + 40))
            roi = (x0, y0, x1, y1)
            if 0.5 * (ow / (x1 - x0) + oh / (y1 - y0)) < 0.08:
                continue
            out = dtt.croprescale_image_cv2(img, torch.tensor(roi, dtype=torch.int32), (ow, oh), downfilter=kind)
            model = geo.croprescale_image(frame, roi, (ow, oh), use_model=True, downfilter=kind)
        else:
            base = Affine2d.range_remap_2d([0.0, 0.0], [float(w), float(h)], [0.0, 0.0], [float(ow), float(oh)])
            tr = Affine2d.trs(translations=torch.tensor([float(rng.uniform(-3, 3)), float(rng.uniform(-3, 3))]),
                              angles=torch.tensor(float(rng.uniform(-0.6, 0.6))), scales=torch.tensor(float(rng.uniform(0.8, 1.3)))) @ base
            if float(tr.scales) < 0.12:
                continue
            out = dtt.affine_transform_image_cv2(img, tr, (ow, oh), downfilter=kind)
            model = geo.affine_transform_image(frame, tr.tensor().numpy(), (ow, oh), use_model=True, downfilter=kind)
        assert np.array_equal(out.cpu().numpy()[0], model), (t, kind)
