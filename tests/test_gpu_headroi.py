"""PutRoiFromLandmarks(extend_to_forehead=True) / roi_override='extent_to_forehead' on the GPU (b200aug_head_roi) against
outputs of the unmodified reference with its real face model (tests/golden/headroi.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "headroi.npz")
TOL = dict(rtol=1e-5, atol=3e-4)  # labels: 1e-4 relative on coordinates of a few hundred pixels


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_mean_shape_roi_matches_reference(gold):
    from trackertraincode_b200.facemodel import HeadModel

    model = HeadModel(gold["hull_vertices"])
    r = model.roi(cuda(gold["coord"]), cuda(gold["pose"])).cpu().numpy()
    np.testing.assert_allclose(r, gold["roi_full"], **TOL)
    # one unbatched sample, and a batch that is not a multiple of the kernel's sample group
    r1 = model.roi(cuda(gold["coord"][3]), cuda(gold["pose"][3])).cpu().numpy()
    assert r1.shape == (4,)
    np.testing.assert_allclose(r1, gold["roi_full"][3], **TOL)
    r13 = model.roi(cuda(gold["coord"][:13]), cuda(gold["pose"][:13])).cpu().numpy()
    assert np.array_equal(r13, r[:13])
    # the half-pixel shift of offset_points_by_half_pixel as an argument
    rs = model.roi(cuda(gold["coord"]), cuda(gold["pose"]), xy_offset=0.5).cpu().numpy()
    np.testing.assert_allclose(rs, gold["roi_full"] + 0.5, **TOL)


def test_deformed_model_matches_reference_module(gold):
    from oracle import headmodel as ohm
    from trackertraincode_b200.facemodel import HeadModel

    model = HeadModel(gold["syn_vertices"], gold["syn_base"])
    r = model.roi(cuda(gold["coord"]), cuda(gold["pose"]), cuda(gold["syn_shape"])).cpu().numpy()
    np.testing.assert_allclose(r, gold["syn_roi"], **TOL)
    np.testing.assert_allclose(r, ohm.head_roi(gold["syn_vertices"], gold["syn_base"], gold["coord"], gold["pose"], gold["syn_shape"]), **TOL)


def test_put_roi_from_landmarks_extend_to_forehead(gold):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import batch as dtb
    from trackertraincode_b200.facemodel import HeadModel

    n = 16
    cats = {"coord": FieldCategory.xys, "pose": FieldCategory.quat, "pt3d_68": FieldCategory.points, "shapeparam": FieldCategory.general}
    data = {"coord": cuda(gold["coord"][:n]), "pose": cuda(gold["pose"][:n]), "pt3d_68": cuda(gold["pt3d_68"][:n]),
            "shapeparam": cuda(gold["syn_shape"][:n])}
    put = dtb.PutRoiFromLandmarks(extend_to_forehead=True, headmodel=HeadModel(gold["hull_vertices"]))
    out = put(Batch(Metadata((450, 450), n, None, None, dict(cats)), dict(data)))
    # the reference ignores "shapeparam" (it looks for a key "shapeparams"): mean shape
    np.testing.assert_allclose(out["roi"].cpu().numpy(), gold["roi_full"][:n], **TOL)
    assert out.meta.categories["roi"] == FieldCategory.roi
    # with that key present the shape parameters are used
    syn = dtb.PutRoiFromLandmarks(extend_to_forehead=True, headmodel=HeadModel(gold["syn_vertices"], gold["syn_base"]))
    data2 = dict(data, shapeparams=cuda(np.zeros((n, 1), np.float32)))
    out2 = syn(Batch(Metadata((450, 450), n, None, None, dict(cats, shapeparams=FieldCategory.general)), data2))
    np.testing.assert_allclose(out2["roi"].cpu().numpy(), gold["syn_roi"][:n], **TOL)
    # no landmarks in the sample: untouched (misc.py:29-30)
    b3 = Batch(Metadata((450, 450), n, None, None, {"coord": FieldCategory.xys}), {"coord": data["coord"]})
    assert "roi" not in put(b3)


def test_fused_augmentation_extent_to_forehead(gold):
    """roi_override='extent_to_forehead' (pipelines.py:352-356) == the 'original' chain fed with the head-model roi."""
    import cases
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation
    from trackertraincode_b200.facemodel import HeadModel

    n = 8
    rng = np.random.default_rng(3)
    frames = np.stack([cases.make_image(rng, 450, 450, "smooth") for _ in range(n)])
    cats = {"image": FieldCategory.image, "roi": FieldCategory.roi, "coord": FieldCategory.xys, "pose": FieldCategory.quat,
            "pt3d_68": FieldCategory.points}
    coord = gold["coord"][:n].copy()
    coord[:, 2] = np.clip(coord[:, 2], 40, 70)
    data = {"image": cuda(frames[:, None]), "roi": cuda(np.tile(np.float32([[100, 100, 300, 300]]), (n, 1))), "coord": cuda(coord),
            "pose": cuda(gold["pose"][:n]), "pt3d_68": cuda(gold["pt3d_68"][:n])}
    model = HeadModel(gold["hull_vertices"])
    head = FusedPoseAugmentation(129, roi_override="extent_to_forehead", train=False, headmodel=model)
    plain = FusedPoseAugmentation(129, roi_override="original", train=False)
    out = head(Batch(Metadata((450, 450), n, None, None, dict(cats)), dict(data)))
    data2 = dict(data, roi=model.roi(data["coord"], data["pose"], xy_offset=0.5))
    want = plain(Batch(Metadata((450, 450), n, None, None, dict(cats)), data2))
    head.status.flush()
    for k in ("image", "roi", "coord", "pt3d_68", "pose"):
        assert torch.equal(out[k], want[k]), k
    # and it is a different crop than the one of the dataset's own roi
    other = plain(Batch(Metadata((450, 450), n, None, None, dict(cats)), dict(data)))
    assert not torch.equal(out["image"], other["image"])
