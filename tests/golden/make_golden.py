#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference) on seeded inputs.

Run in the authoring container only:  python tests/golden/make_golden.py
The reference needs h5py / kornia / strenum / matplotlib at import time; inert stand-ins live in
tests/golden/_stubs (none of their code runs on the geometric/label half exercised here).

Outputs (all arrays indexed by case number of tests/golden/cases.py):
  focus_chain.npz   per-stage outputs of  offset_points_by_half_pixel -> RandomFocusRoi(explicit params)
                    -> horizontal_flip_and_rot_90(explicit draws) -> normalize_batch
  algebra.npz       Affine2d / apply_affine2d / torchquaternion known-answer vectors on random inputs
  env.json          versions of torch / cv2 / numpy that produced them
"""
import json
import os
import sys
from functools import partial
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", HERE]

import kornia_stub  # noqa: E402

kornia_stub.install()

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import trackertraincode.datatransformation as dtr  # noqa: E402
from trackertraincode.datasets.batch import Batch, Metadata  # noqa: E402
from trackertraincode.datasets.dshdf5pose import FieldCategory  # noqa: E402
from trackertraincode.datatransformation.batch.geometric import GeneralFocusRoi  # noqa: E402
from trackertraincode.datatransformation.tensors.affinetrafo import (  # noqa: E402
    apply_affine2d, transform_coord, transform_keypoints, transform_roi, transform_rot)
from trackertraincode.facemodel.keypoints68 import flip_map  # noqa: E402
from trackertraincode.neuralnets import torchquaternion as tq  # noqa: E402
from trackertraincode.neuralnets.affine2d import Affine2d  # noqa: E402

import cases  # noqa: E402

CATS = dict(image=FieldCategory.image, roi=FieldCategory.roi, coord=FieldCategory.xys, pose=FieldCategory.quat,
            pt3d_68=FieldCategory.points, shapeparam=FieldCategory.general)
LABELS = ("roi", "coord", "pose", "pt3d_68", "shapeparam")


def to_batch(c) -> Batch:
    w, h = c["wh"]
    data = {"image": torch.from_numpy(c["image"][..., None].copy())}
    for k in LABELS:
        data[k] = torch.from_numpy(c[k].copy())
    return Batch(Metadata((w, h), 0, "golden", None, categories=dict(CATS)), data)


def run_case(c):
    S = c["out_size"]
    out = {}
    sample = dtr.batch.offset_points_by_half_pixel(to_batch(c))
    params = dtr.batch.RoiFocusRandomizationParameters(
        scales=torch.tensor(float(c["scale"]), dtype=torch.float32),
        angles=torch.tensor(float(c["angle"]), dtype=torch.float32),
        translations=torch.from_numpy(c["translation"].copy()),
        upfilter="linear", downfilter="area")
    focus = dtr.batch.RandomFocusRoi(S)
    focus.make_randomization_parameters = lambda B: params
    # intermediates, through the reference's own helpers
    view = GeneralFocusRoi._compute_view_roi(sample["roi"], params.scales, params.translations, 0.3)
    view_i = torch.round(view).to(torch.int32)
    tr = focus._center_rotation_tr(params.angles) @ focus._compute_point_transform_from_roi((), view_i, S)
    out["view_roi"] = view_i.numpy()
    out["tr"] = tr.tensor().numpy()
    sample = focus(sample)
    out["focus_image"] = sample["image"].numpy()[0]
    for k in LABELS:
        out["focus_" + k] = sample[k].numpy()
    # flip / rot90 with the two numpy draws forced
    with mock.patch.object(np.random, "randint", lambda a, b: 0 if c["do_flip"] else 1), \
            mock.patch.object(np.random, "choice", lambda vals, p=None: c["rot_dir"]):
        sample = dtr.batch.horizontal_flip_and_rot_90(0.01, sample)
    out["flip_image"] = np.ascontiguousarray(sample["image"].numpy()[0])
    for k in LABELS:
        out["flip_" + k] = sample[k].numpy()
    sample = dtr.batch.normalize_batch(sample)
    out["final_image"] = np.ascontiguousarray(sample["image"].numpy()[0])
    for k in LABELS:
        out["final_" + k] = sample[k].numpy()
    out["final_image_whitened"] = dtr.tensors.whiten_image(sample["image"]).numpy()[0]
    return out


def make_focus_chain():
    rows = [run_case(cases.make_case(i)) for i in range(cases.N_CASES)]
    packed = {k: np.stack([r[k] for r in rows], 0) for k in rows[0]}
    # normalised images are u8/256 exactly; keep them out of the file (recomputed in the test), keep u8 crops
    del packed["final_image"], packed["final_image_whitened"]
    np.savez_compressed(os.path.join(HERE, "focus_chain.npz"), **packed)
    print("focus_chain.npz:", {k: v.shape for k, v in packed.items()})


def make_algebra():
    g = torch.Generator().manual_seed(7)
    n = 64
    ang = (torch.rand(n, generator=g) * 2 - 1) * 3.0
    sc = torch.rand(n, generator=g) * 1.5 + 0.3
    t = torch.randn(n, 2, generator=g) * 50
    out = {}
    trs, mats, invs, scs, dets, prods = [], [], [], [], [], []
    pts_o, roi_o, coord_o, quat_o, bt_o = [], [], [], [], []
    pts = torch.randn(n, 68, 3, generator=g) * 40 + 100
    roi = torch.randn(n, 4, generator=g) * 30 + torch.tensor([80.0, 90.0, 200.0, 210.0])
    coord = torch.randn(n, 3, generator=g) * 20 + 100
    quat = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=-1)
    for i in range(n):
        a = Affine2d.trs(translations=t[i], angles=ang[i], scales=sc[i])
        if i % 3 == 0:  # mirrored maps exercise flip_map and the quaternion reflection
            a = a @ Affine2d.range_remap_2d([0.0, 0.0], [129, 129], [129, 0], [0, 129])
        b = Affine2d.trs(translations=t[(i + 1) % n], angles=ang[(i + 5) % n], scales=sc[(i + 3) % n])
        mats.append(a.tensor().numpy())
        prods.append((a @ b).tensor().numpy())
        invs.append(a.inv().tensor().numpy())
        scs.append(a.scales.numpy())
        dets.append(a.det.numpy())
        pts_o.append(transform_keypoints(a, pts[i]).numpy())
        roi_o.append(transform_roi(a, roi[i]).numpy())
        coord_o.append(transform_coord(a, coord[i]).numpy())
        quat_o.append(transform_rot(a, quat[i]).numpy())
        bt_o.append(apply_affine2d(a, "image_backtransform", b.tensor(), FieldCategory.general).numpy())
    out = dict(angles=ang.numpy(), scales_in=sc.numpy(), translations=t.numpy(), pts=pts.numpy(), roi=roi.numpy(),
               coord=coord.numpy(), quat=quat.numpy(), mats=np.stack(mats), prods=np.stack(prods), invs=np.stack(invs),
               scales=np.stack(scs), dets=np.stack(dets), pts_out=np.stack(pts_o), roi_out=np.stack(roi_o),
               coord_out=np.stack(coord_o), quat_out=np.stack(quat_o), backtransform_out=np.stack(bt_o),
               flip_map=np.asarray(flip_map, np.int64),
               quat_mult=tq.mult(quat, quat.roll(1, 0)).numpy(), quat_matrix=tq.tomatrix(quat).numpy())
    np.savez_compressed(os.path.join(HERE, "algebra.npz"), **out)
    print("algebra.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    torch.set_num_threads(1)
    make_focus_chain()
    make_algebra()
    with open(os.path.join(HERE, "env.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cv2=cv2.__version__, numpy=np.__version__,
                       reference_commit="8d2478c4", generator="tests/golden/make_golden.py"), f, indent=1)
