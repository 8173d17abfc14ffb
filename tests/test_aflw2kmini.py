"""BASELINE.json configs[0]: the reference's bundled fixture `aflw2kmini.h5` (16 JPEG frames 450 x 450 + labels) through
the augmentation chain, batch 64 = the 16 samples under 4 parameter draws, 129 x 129 gray crops.  Golden outputs come from the
UNMODIFIED reference (tests/golden/make_golden_aflw2kmini.py; the HDF5 file is read with tests/golden/minihdf5.py).
CPU: the oracle reproduces them.  GPU: the fused kernel reproduces them (uint8 crops bit-exact, labels 1e-4), from frames
decoded by cv2 like the reference does, and -- within the JPEG decoders' rounding -- from frames decoded on the GPU."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import geometric as ogeo, normalization as onrm
from oracle.geometric import Sample

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "aflw2kmini.npz"))
S, N, B = 129, 16, 64
CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts", shapeparam="")
LABELS = ("roi", "coord", "pose", "pt3d_68", "shapeparam")


def blobs():
    o = G["jpeg_offsets"]
    return [G["jpeg_bytes"][o[i]:o[i + 1]] for i in range(N)]


def host_cos_sin(a):
    t = torch.from_numpy(np.asarray(a, np.float32))
    return torch.cos(t).numpy(), torch.sin(t).numpy()


def check_labels(got, j, quat_tol=2e-5, rtol=1e-4, atol=2e-5):
    for k in ("roi", "coord", "pt3d_68", "shapeparam"):
        np.testing.assert_allclose(got[k], G["final_" + k][j], rtol=rtol, atol=atol, err_msg=f"{k} of sample {j}")
    q, w = got["pose"], G["final_pose"][j]
    assert min(np.abs(q - w).max(), np.abs(q + w).max()) < quat_tol, f"pose of sample {j}"


def test_oracle_reproduces_the_reference_on_the_bundled_fixture():
    frames = [cv2.imdecode(b, 0) for b in blobs()]
    cs, sn = host_cos_sin(G["angles"])
    for j in range(B):
        i = j % N
        data = {"image": frames[i][..., None], **{k: G["in_" + k][i] for k in LABELS}}
        s = onrm.offset_points_by_half_pixel(Sample((450, 450), data, dict(CATS)))
        s, inter = ogeo.focus_roi(s, ogeo.RoiFocusParams(G["scales"][j], G["angles"][j], G["translations"][j], (float(cs[j]), float(sn[j]))), S)
        assert np.array_equal(inter["view_roi"], G["view_roi"][j])
        s = ogeo.horizontal_flip_and_rot_90(s, bool(G["do_flip"][j]), int(G["rot_dir"][j]))
        assert np.array_equal(s.data["image"].reshape(S, S), G["flip_image"][j]), f"crop of sample {j}"
        check_labels(onrm.normalize_sample(s).data, j, quat_tol=1e-5, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("decoder", ["cv2", "nvjpeg"])
def test_fused_kernel_on_the_bundled_fixture(decoder):
    from trackertraincode_b200 import _native as Nn
    from trackertraincode_b200.datasets import preprocessing as pre
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import _engine as E

    if decoder == "cv2":
        frames = torch.from_numpy(np.stack([cv2.imdecode(b, 0) for b in blobs()])).cuda()
    else:
        frames = pre.imdecode_batch(blobs(), stack=True)
    idx = torch.arange(B) % N
    data = {"image": frames[idx.cuda()]}
    for k in LABELS:
        data[k] = torch.from_numpy(G["in_" + k])[idx].cuda()
    batch = Batch(Metadata((450, 450), B, "aflw2kmini", None, {k: FieldCategory(v) for k, v in CATS.items()}), data)
    an = torch.from_numpy(G["angles"])
    geo = E.GeoParams(torch.from_numpy(G["scales"]), an, torch.from_numpy(G["translations"]), E.host_cos_sin(an))
    r = E.fused_forward(batch, flags=Nn.F_HALF_PIXEL | Nn.F_FOCUS | Nn.F_FLIPROT | Nn.F_NORMALIZE, out_size=S, geo=geo,
                        do_flip=torch.from_numpy(G["do_flip"].astype(np.uint8)), rot_dir=torch.from_numpy(G["rot_dir"]),
                        want_view_roi=True, want_status=True)
    assert not r.status.cpu().numpy().any()
    assert np.array_equal(r.view_roi.cpu().numpy(), G["view_roi"])
    img = r.batch["image"].cpu().numpy()[:, 0] * 256.0
    if decoder == "cv2":
        assert np.array_equal(img, G["flip_image"].astype(np.float32)), "uint8 crops differ from the reference's"
    else:  # the decoders differ by at most 2 grey levels per source pixel; area averaging cannot enlarge that
        assert np.abs(img - G["flip_image"]).max() <= 2.0
    for j in range(B):
        check_labels({k: r.batch[k][j].cpu().numpy() for k in LABELS}, j)
