#!/usr/bin/env python
"""Generate tests/golden/upfilter.npz: the UNMODIFIED reference (from /root/reference) run with the up-filters 'cubic' and
'lanczos' (trackertraincode/datatransformation/tensors/image_geometric_cv2.py:65-82, 105-119).

Run in the authoring container only:  python tests/golden/make_golden_upfilter.py

  focus_<filter>    RandomFocusRoi(129) behind offset_points_by_half_pixel on FOCUS_CASES of tests/golden/cases.py with
                    `upfilter=<filter>` in the randomization parameters (batch/geometric.py:193-231)
  tensor_<filter>_j the tensor-level entries on TENSOR_CASES; tensor_tr: the float32 transforms of the 'affine' entries
  ipp               whether cv2 ran with Intel IPP (it replaces OpenCV's INTER_CUBIC resize kernel; see oracle/cv2_model.py)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", HERE]

import kornia_stub  # noqa: E402

kornia_stub.install()

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import trackertraincode.datatransformation as dtr  # noqa: E402
from trackertraincode.datatransformation.tensors.image_geometric_cv2 import (  # noqa: E402
    affine_transform_image_cv2, croprescale_image_cv2)
from trackertraincode.neuralnets.affine2d import Affine2d  # noqa: E402

import cases  # noqa: E402
from make_golden import to_batch  # noqa: E402
from upfilter_cases import FILTERS, FOCUS_CASES, TENSOR_CASES  # noqa: E402


def tensor_transform(c, out_wh, g):
    w, h = c["wh"]
    ow, oh = out_wh
    angle, scale, tx, ty = g
    base = Affine2d.range_remap_2d([0.0, 0.0], [float(w), float(h)], [0.0, 0.0], [float(ow), float(oh)])
    return Affine2d.trs(translations=torch.tensor([tx, ty]), angles=torch.tensor(angle), scales=torch.tensor(scale)) @ base


def main():
    torch.set_num_threads(1)
    out = {}
    for f in FILTERS:
        rows = []
        for i in FOCUS_CASES:
            c = cases.make_case(i)
            sample = dtr.batch.offset_points_by_half_pixel(to_batch(c))
            params = dtr.batch.RoiFocusRandomizationParameters(
                scales=torch.tensor(float(c["scale"]), dtype=torch.float32),
                angles=torch.tensor(float(c["angle"]), dtype=torch.float32),
                translations=torch.from_numpy(c["translation"].copy()), upfilter=f, downfilter="area")
            focus = dtr.batch.RandomFocusRoi(c["out_size"])
            focus.make_randomization_parameters = lambda B, params=params: params
            rows.append(focus(sample)["image"].numpy()[0])
        out["focus_" + f] = np.stack(rows, 0)
        for j, (i, entry, out_wh, g) in enumerate(TENSOR_CASES):
            c = cases.make_case(i)
            img = torch.from_numpy(c["image"][None].copy())  # [1, H, W]
            if entry == "crop":
                res = croprescale_image_cv2(img, torch.tensor(g, dtype=torch.int32), out_wh, upfilter=f)
            else:
                res = affine_transform_image_cv2(img, tensor_transform(c, out_wh, g), out_wh, upfilter=f)
            out[f"tensor_{f}_{j}"] = res.numpy()[0]
    trs = np.zeros((len(TENSOR_CASES), 2, 3), np.float32)
    for j, (i, entry, out_wh, g) in enumerate(TENSOR_CASES):
        if entry == "affine":
            trs[j] = tensor_transform(cases.make_case(i), out_wh, g).tensor().numpy()
    out["tensor_tr"] = trs
    out["ipp"] = np.asarray(bool(cv2.ipp.useIPP()))
    np.savez_compressed(os.path.join(HERE, "upfilter.npz"), **out)
    print({k: v.shape for k, v in out.items()}, "ipp:", out["ipp"])


if __name__ == "__main__":
    main()
