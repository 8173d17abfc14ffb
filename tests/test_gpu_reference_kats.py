"""The reference's own image/label consistency tests for this path (test/test_affine_img_trafo.py:95-256), run through the
B200 mirror API on the GPU: a heat map with three peaks is cropped / rescaled / rotated together with its landmark labels
and the peaks' centroids must land on the transformed landmarks.  The reference uses a 3-channel float heat map; this path
is single-channel uint8, so every channel is one call with identical parameters (same arithmetic per channel in cv2).
All six rows of the reference's filter table (:175-192) run: linear / cubic / lanczos up, gaussian / hamming / area down
(SURVEY.md 8 a9)."""
from functools import partial

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_test_data(way):  # test/test_affine_img_trafo.py:100-142
    if way == "down":
        S, R = 200, 20
        points = torch.tensor([[15, 15, 0], [45, 35, 0], [85, 85, 0]], dtype=torch.float32) + 50
    else:
        S, R = 20, 200
        points = torch.tensor([[1, 1, 0], [4, 3, 0], [8, 8, 0]], dtype=torch.float32) + 0.5 + 5
    img = torch.zeros((3, 20, 20), dtype=torch.uint8)
    img[0, 5 + 1, 5 + 1] = 255
    img[1, 5 + 3, 5 + 4] = 255
    img[2, 5 + 8, 5 + 8] = 255
    if way == "down":
        img = img.repeat_interleave(10, dim=1).repeat_interleave(10, dim=2)
    return S, R, img, points, torch.tensor([0.0, 0.0, S, S])


def centroids(hm):  # kornia.spatial_expectation2d(normalized_coordinates=False) + 0.5, test_affine_img_trafo.py:145-155
    hm = hm.double()
    hm = hm / hm.sum(dim=(-1, -2), keepdim=True)
    ys, xs = torch.meshgrid(torch.arange(hm.shape[-2], dtype=torch.float64), torch.arange(hm.shape[-1], dtype=torch.float64), indexing="ij")
    return torch.stack([(hm * xs).sum((-1, -2)), (hm * ys).sum((-1, -2))], -1) + 0.5


def run_focus(way, params_fn):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import batch as dtb

    S, R, img, points, roi = make_test_data(way)
    outs = []
    for c in range(3):
        b = Batch(Metadata((S, S), 0, None, None, {"image": FieldCategory.image, "pt3d_68": FieldCategory.points, "roi": FieldCategory.roi}),
                  {"image": img[c][..., None].cuda(), "pt3d_68": points.cuda(), "roi": roi.cuda()})
        aug = dtb.RandomFocusRoi(new_size=R)
        aug.make_randomization_parameters = params_fn
        outs.append(aug(b))
    image = torch.stack([o["image"].reshape(R, R).cpu() for o in outs])
    return R, image, outs[0]["pt3d_68"].cpu()


def no_randomization(B, filter_args):  # :73-79
    from trackertraincode_b200.datatransformation import batch as dtb
    return dtb.RoiFocusRandomizationParameters(scales=torch.tensor(1.0), angles=torch.tensor(0.0), translations=torch.tensor([0.0, 0.0]), **filter_args)


def with_some_similarity_trafo(B, filter_args):  # :82-88
    from trackertraincode_b200.datatransformation import batch as dtb
    return dtb.RoiFocusRandomizationParameters(scales=torch.tensor(0.75), angles=torch.tensor(20.0 * np.pi / 180.0),
                                               translations=torch.tensor([-0.1, 0.03]), **filter_args)


CONFIGS = [("up", {"upfilter": "linear"}, 0.0), ("up", {"upfilter": "cubic"}, 0.5), ("up", {"upfilter": "lanczos"}, 0.5),
           ("down", {"downfilter": "gaussian"}, 0.0), ("down", {"downfilter": "hamming"}, 0.0), ("down", {"downfilter": "area"}, 0.0)]  # :175-192


@pytest.mark.parametrize("way, filter_args, tol", CONFIGS)
def test_scalingtrafo(way, filter_args, tol):  # :195-203
    R, image, pts = run_focus(way, partial(no_randomization, filter_args=filter_args))
    assert image.shape == (3, R, R)
    np.testing.assert_allclose(centroids(image).numpy(), pts[:, :2].numpy(), atol=0.01 + tol)


@pytest.mark.parametrize("way, filter_args, tol", CONFIGS)
def test_scalingtrafo_with_randomizer(way, filter_args, tol):  # :206-216
    R, image, pts = run_focus(way, partial(with_some_similarity_trafo, filter_args=filter_args))
    np.testing.assert_allclose(centroids(image).numpy(), pts[:, :2].numpy(), atol=0.01 + tol + (0.2 if way == "down" else 0.5))


@pytest.mark.parametrize("filter_args, exc", [({"upfilter": "nearest"}, KeyError), ({"downfilter": "box"}, NotImplementedError)])
def test_unknown_filters_raise_like_the_reference(filter_args, exc):  # image_geometric_cv2.py:62, 70-75
    with pytest.raises(exc):
        run_focus("up", partial(no_randomization, filter_args=filter_args))


def test_against_the_oracle_on_the_same_fixture():
    """The same fixture through the CPU oracle (cv2): pixels bit-exact, i.e. the consistency above is the reference's own."""
    from oracle import geometric as ogeo
    from oracle.geometric import Sample

    for way, (scale, angle, tr) in (("down", (0.75, 20.0 * np.pi / 180.0, (-0.1, 0.03))), ("up", (1.0, 0.0, (0.0, 0.0)))):
        S, R, img, points, roi = make_test_data(way)
        fn = partial(with_some_similarity_trafo if way == "down" else no_randomization, filter_args={})
        _, image, pts = run_focus(way, fn)
        a32 = np.float32(angle)
        for c in range(3):
            s = Sample((S, S), {"image": img[c].numpy()[..., None], "pt3d_68": points.numpy(), "roi": roi.numpy()},
                       {"image": "img", "pt3d_68": "pts", "roi": "roi"})
            want, _ = ogeo.focus_roi(s, ogeo.RoiFocusParams(np.float32(scale), a32, np.asarray(tr, np.float32),
                                                            (float(torch.cos(torch.tensor(a32))), float(torch.sin(torch.tensor(a32))))), R)
            assert np.array_equal(image[c].numpy(), want.data["image"].reshape(R, R)), (way, c)
        np.testing.assert_allclose(pts.numpy(), want.data["pt3d_68"], rtol=1e-4, atol=2e-5)
