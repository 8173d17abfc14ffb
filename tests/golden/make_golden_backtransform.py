#!/usr/bin/env python
"""Generate tests/golden/backtransform.npz by running the UNMODIFIED reference's evaluation-side chain
(eval.py:149-212, `Predictor.predict_batch`):

    FocusRoi(129, 1.1, insert_backtransform=True)  per frame  ->  Batch.collate  ->  normalize_batch
      ->  [network]  ->  unnormalize_batch  ->  _apply_backtrafo(Affine2d(image_backtransform))

with a stand-in network that returns seeded "predictions" in the normalised frame, plus two chains that exercise the
`image_backtransform` rewriting of apply_affine2d (tensors/affinetrafo.py:137-147) outside that cancelling pair:
FocusRoi(insert) -> horizontal_flip_and_rot_90 (forced draws), and a second FocusRoi WITHOUT insert on a sample that
already carries a back-transform.

Run in the authoring container only:  python tests/golden/make_golden_backtransform.py
Inputs are regenerated from tests/golden/cases.py (case i, the 450x450 / 320x240 / ... frames); only outputs are stored.
"""
import os
import sys
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", HERE]

import kornia_stub  # noqa: E402

kornia_stub.install()

import numpy as np  # noqa: E402
import torch  # noqa: E402

import trackertraincode.datatransformation as dtr  # noqa: E402
from trackertraincode import eval as ref_eval  # noqa: E402
from trackertraincode.datasets.batch import Batch, Metadata  # noqa: E402
from trackertraincode.neuralnets.affine2d import Affine2d  # noqa: E402

import cases  # noqa: E402

S = 129
CASES = [0, 1, 2, 3, 4, 5, 6, 7, 30, 31, 33]  # all frame sizes, a tiny face (up-scaling), a box over the border


def fake_predictions(n: int):
    """What a pose network would return for n crops, in the normalised frame (seeded)."""
    rng = np.random.default_rng(77)
    q = rng.standard_normal((n, 4))
    return dict(coord=np.concatenate([rng.uniform(-0.6, 0.6, (n, 2)), rng.uniform(0.2, 0.8, (n, 1))], -1).astype(np.float32),
                pose=(q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32),
                pt3d_68=rng.uniform(-0.9, 0.9, (n, 68, 3)).astype(np.float32),
                roi=np.sort(rng.uniform(-0.8, 0.8, (n, 2, 2)), axis=1).reshape(n, 4).astype(np.float32))


class StandInNet(ref_eval.InferenceNetwork):
    def __init__(self, preds):
        self.preds = {k: torch.from_numpy(v) for k, v in preds.items()}
        self.seen = None

    def __call__(self, images):
        self.seen = images.clone()
        return {k: v.clone() for k, v in self.preds.items() if k != "roi"}

    @property
    def input_resolution(self):
        return S

    @property
    def device_for_input(self):
        return "cpu"


def main():
    cs = [cases.make_case(i) for i in CASES]
    n = len(cs)
    preds = fake_predictions(n)
    out = {}
    # ---- the reference's own predict_batch (images [H,W,C] uint8, one Predictor per frame size class is not needed:
    # Predictor builds per-frame samples).  Frames of different sizes cannot be collated by the reference (meta differs), so
    # the chain runs per frame-size group; results are stored per case.
    final = {k: np.zeros_like(preds[k]) for k in ("coord", "pose", "pt3d_68")}
    focus_bt = np.zeros((n, 2, 3), np.float32)
    norm_bt = np.zeros((n, 2, 3), np.float32)
    unnorm_bt = np.zeros((n, 2, 3), np.float32)
    net_in = np.zeros((n, 1, S, S), np.float32)
    unnorm = {k: np.zeros_like(preds[k]) for k in ("coord", "pose", "pt3d_68")}
    sizes = sorted({c["wh"] for c in cs})
    for wh in sizes:
        idx = [i for i, c in enumerate(cs) if c["wh"] == wh]
        images = [torch.from_numpy(cs[i]["image"][..., None].copy()) for i in idx]
        rois = torch.from_numpy(np.stack([cs[i]["roi"] for i in idx]))
        net = StandInNet({k: v[idx] for k, v in preds.items()})
        p = ref_eval.Predictor(net, focus_roi_expansion_factor=1.1)
        res = p.predict_batch(images, rois)
        for k in final:
            final[k][idx] = res[k].numpy()
        net_in[idx] = net.seen.numpy()
        # the same chain step by step, for the intermediates
        batch = Batch.collate([p._create_sample(i, r) for i, r in zip(images, rois)])
        focus_bt[idx] = batch["image_backtransform"].numpy()
        batch = dtr.batch.normalize_batch(batch)
        norm_bt[idx] = batch["image_backtransform"].numpy()
        pb = Batch(batch.meta, **{k: torch.from_numpy(preds[k][idx].copy()) for k in ("coord", "pose", "pt3d_68")})
        pb.meta.categories.update({"coord": dtr.FieldCategory.xys, "pose": dtr.FieldCategory.quat, "pt3d_68": dtr.FieldCategory.points})
        pb["image_backtransform"] = batch["image_backtransform"]
        pb = dtr.batch.unnormalize_batch(pb)
        unnorm_bt[idx] = pb["image_backtransform"].numpy()
        for k in unnorm:
            unnorm[k][idx] = pb[k].numpy()
        chk = ref_eval._apply_backtrafo(Affine2d(pb.pop("image_backtransform")), pb)
        for k in final:  # the step-by-step chain IS predict_batch
            assert np.array_equal(chk[k].numpy(), final[k][idx]), k
    out.update(focus_bt=focus_bt, norm_bt=norm_bt, unnorm_bt=unnorm_bt, net_in=net_in)
    out.update({"final_" + k: v for k, v in final.items()})
    out.update({"unnorm_" + k: v for k, v in unnorm.items()})

    # ---- FocusRoi(insert) -> flip / rot90 with forced draws: BT @ tr_flip^-1 (no cancelling partner)
    flip_bt = np.zeros((n, 2, 3), np.float32)
    again_bt = np.zeros((n, 2, 3), np.float32)
    again_roi = np.zeros((n, 4), np.float32)
    for j, c in enumerate(cs):
        w, h = c["wh"]
        sample = Batch.from_data_with_categories(Metadata((w, h), 0), {
            "image": (torch.from_numpy(c["image"][None].copy()), dtr.FieldCategory.image),
            "roi": (torch.from_numpy(c["roi"].copy()), dtr.FieldCategory.roi)})
        sample = dtr.batch.FocusRoi(S, 1.1, insert_backtransform=True)(sample)
        do_flip, rot_dir = bool(j % 2 == 0), int([0, 1, -1][j % 3])
        with mock.patch.object(np.random, "randint", lambda a, b: 0 if do_flip else 1), \
                mock.patch.object(np.random, "choice", lambda vals, p=None: rot_dir):
            flipped = dtr.batch.horizontal_flip_and_rot_90(0.01, sample)
        flip_bt[j] = flipped["image_backtransform"].numpy()
        # ---- a second focus WITHOUT insert_backtransform on the sample that carries one: BT @ tr2^-1
        second = dtr.batch.FocusRoi(65, 1.3, insert_backtransform=False)(flipped)
        again_bt[j] = second["image_backtransform"].numpy()
        again_roi[j] = second["roi"].numpy()
    out.update(flip_bt=flip_bt, again_bt=again_bt, again_roi=again_roi,
               flip_draws=np.asarray([[int(j % 2 == 0), [0, 1, -1][j % 3]] for j in range(n)], np.int32))
    out["cases"] = np.asarray(CASES, np.int32)
    for k, v in preds.items():
        out["pred_" + k] = v
    np.savez_compressed(os.path.join(HERE, "backtransform.npz"), **out)
    print("wrote backtransform.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
