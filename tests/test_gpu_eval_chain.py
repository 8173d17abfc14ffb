"""GPU parity of the evaluation-side surface (through the C ABI, via the mirror API):
  * the Predictor chain of eval.py:149-212 -- FocusRoi(insert_backtransform=True) -> normalize_batch -> [net] ->
    unnormalize_batch -> _apply_backtrafo -- against golden vectors from the UNMODIFIED reference
    (tests/golden/make_golden_backtransform.py), including every intermediate value of `image_backtransform`;
  * `image_backtransform` through flip / rot90 and through a second focus without insert_backtransform
    (tensors/affinetrafo.py:137-147);
  * the tensor-level entries of datatransformation.tensors (croprescale_image_cv2, affine_transform_image_cv2,
    apply_affine2d, whiten/unwhiten, ensure_image_*), batch.to_numpy / to_tensor;
  * video batches (meta.seq): one crop draw per clip (geometric.py:180-191);
  * bool fields -> 0.1 / 0.9 label smoothing inside FusedPoseAugmentation (normalization.py:26-30);
  * RandomGaussianNoiseWithClipping (batch/intensity.py:43-53).
Bars: integer work bit-exact, pixels within 1/255 (bit-exact on uint8), labels 1e-4 relative."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import affine, geometric as ogeo, normalization as onrm, photometric as opho, pipeline as opipe
from oracle.geometric import Sample

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "backtransform.npz"))
S = 129
PRED_CATS = dict(coord="xys", pose="q", pt3d_68="pts")


def _close(got, want, rtol=1e-4, atol=1e-3, msg=""):
    np.testing.assert_allclose(got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else got, want, rtol=rtol, atol=atol, err_msg=msg)


def _quat_close(got, want, tol=2e-5):
    g = got.detach().cpu().numpy()
    assert np.minimum(np.abs(g - want).max(-1), np.abs(g + want).max(-1)).max() < tol


def _frames(cs):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

    data = {"image": [torch.from_numpy(c["image"][..., None].copy()).cuda() for c in cs],
            "roi": torch.from_numpy(np.stack([c["roi"] for c in cs])).cuda()}
    return Batch(Metadata(cs[0]["wh"], len(cs), "eval", None, {"image": FieldCategory.image, "roi": FieldCategory.roi}), data)


def test_predictor_chain_against_reference_goldens():
    import trackertraincode_b200.datatransformation as dtr
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    cs = [cases.make_case(int(i)) for i in GOLD["cases"]]
    n = len(cs)
    batch = dtr.batch.FocusRoi(S, 1.1, insert_backtransform=True)(_frames(cs))  # ragged frames, one launch
    _close(batch["image_backtransform"], GOLD["focus_bt"], 1e-5, 1e-4, "BT after focus")
    assert tuple(batch["image_original_size"].shape) == (n, 2)
    batch = dtr.batch.normalize_batch(batch)
    _close(batch["image_backtransform"], GOLD["norm_bt"], 1e-5, 1e-4, "BT after normalize")
    net_in = dtr.tensors.whiten_image(batch["image"])
    assert np.array_equal(net_in.cpu().numpy(), GOLD["net_in"]), "the crops the network is fed differ from the reference's"
    # the network's answer (stand-in, the same numbers the reference chain was given), then back to the original frame
    preds = Batch(batch.meta, {k: torch.from_numpy(GOLD["pred_" + k]).cuda() for k in PRED_CATS})
    preds.meta.categories.update({"coord": FieldCategory.xys, "pose": FieldCategory.quat, "pt3d_68": FieldCategory.points})
    preds["image_backtransform"] = batch["image_backtransform"]
    preds = dtr.batch.unnormalize_batch(preds)
    _close(preds["image_backtransform"], GOLD["unnorm_bt"], 1e-5, 1e-4, "BT after unnormalize")
    for k in PRED_CATS:
        (_quat_close if k == "pose" else _close)(preds[k], GOLD["unnorm_" + k])
    bt = Affine2d(preds.pop("image_backtransform"))
    for k, v in list(preds.items()):  # eval.py:149-155 _apply_backtrafo
        out = dtr.tensors.apply_affine2d(bt, k, v, preds.get_category(k))
        (_quat_close if k == "pose" else _close)(out, GOLD["final_" + k])


def test_backtransform_through_flip_and_second_focus():
    import trackertraincode_b200.datatransformation as dtr

    cs = [cases.make_case(int(i)) for i in GOLD["cases"]]
    batch = dtr.batch.FocusRoi(S, 1.1, insert_backtransform=True)(_frames(cs))
    draws = GOLD["flip_draws"]
    flipped = dtr.batch.horizontal_flip_and_rot_90(0.01, batch, draws=(torch.from_numpy(draws[:, 0].astype(np.uint8)),
                                                                      torch.from_numpy(draws[:, 1].astype(np.int8))))
    _close(flipped["image_backtransform"], GOLD["flip_bt"], 1e-5, 1e-4, "BT after flip / rot90")
    second = dtr.batch.FocusRoi(65, 1.3, insert_backtransform=False)(flipped)
    _close(second["roi"], GOLD["again_roi"], 1e-5, 1e-4)
    _close(second["image_backtransform"], GOLD["again_bt"], 1e-4, 1e-3, "BT after a focus without insert_backtransform")


def test_tensor_level_image_entries_vs_oracle():
    """croprescale_image_cv2 / affine_transform_image_cv2 (tensors/image_geometric_cv2.py:85-155) with explicit geometry:
    bit-exact against the same OpenCV calls."""
    import trackertraincode_b200.datatransformation as dtr
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    rng = np.random.default_rng(5)
    for t in range(6):
        h, w = [(450, 450), (240, 320), (180, 200)][t % 3]
        img = cases.make_image(rng, w, h, "noise" if t % 2 else "smooth")
        g = torch.from_numpy(img[None].copy()).cuda()  # [1, H, W]
        # crop boxes inside, over the border, smaller than the output (up-scaling)
        for roi, size in (([20, 30, 20 + 180, 30 + 170], 129), ([-25, 10, 210, 260], (96, 64)), ([50, 60, 110, 115], 129)):
            want = ogeo.croprescale_image(img, np.asarray(roi, np.int32), (size, size) if isinstance(size, int) else size)
            got = dtr.tensors.croprescale_image_cv2(g, torch.tensor(roi, dtype=torch.int32), size)
            assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy()[0], want), (t, roi)
        for ang, sc, size in ((20.0, 0.6, 129), (-33.0, 0.45, (96, 64)), (10.0, 1.7, 129)):
            tr = affine.compose(affine.trs(translations=np.float32([3.0, -7.0]), angles=np.float32(ang * np.pi / 180), scales=np.float32(sc)),
                                affine.range_remap_2d([0, 0], [w, h], [0, 0], [w * 0.4, h * 0.4]))
            wh = (size, size) if isinstance(size, int) else size
            want = ogeo.affine_transform_image(img, tr, wh)
            got = dtr.tensors.affine_transform_image_cv2(g, Affine2d(torch.from_numpy(np.asarray(tr, np.float32))), size)
            assert np.array_equal(got.cpu().numpy()[0], want), (t, ang, sc)
    # layouts: [H, W, C] in, stacks of images, NCHW out
    hwc = torch.from_numpy(img[..., None].copy()).cuda()
    a = dtr.tensors.croprescale_image_cv2(hwc, torch.tensor([5, 5, 150, 160]), 64)
    b = dtr.tensors.croprescale_image_cv2(torch.stack([g, g]), torch.tensor([5, 5, 150, 160]), 64)
    assert a.shape == (1, 64, 64) and b.shape == (2, 1, 64, 64) and torch.equal(b[0], a) and torch.equal(b[1], a)
    assert dtr.tensors.ensure_image_nhwc(a).shape == (64, 64, 1) and dtr.tensors.ensure_image_nchw(hwc).shape == (1, h, w)
    x = torch.rand(2, 1, 8, 8, device="cuda")
    assert torch.equal(dtr.tensors.unwhiten_image(dtr.tensors.whiten_image(x)), x.sub(0.5).add(0.5))
    from trackertraincode_b200 import _native as N_

    with pytest.raises(NotImplementedError):  # image_geometric_cv2.py:62 (the filters themselves: tests/test_gpu_prefilter.py)
        dtr.tensors.croprescale_image_cv2(g, torch.tensor([5, 5, 150, 160]), 64, downfilter="box")
    with pytest.raises(N_.NativeError):  # there is no CPU path
        dtr.tensors.croprescale_image_cv2(g.cpu(), torch.tensor([5, 5, 150, 160]), 64)


def test_to_numpy_to_tensor_roundtrip():
    import trackertraincode_b200.datatransformation as dtr

    cs = [cases.make_case(i) for i in (0, 1)]
    b = _frames(cs)
    nb = dtr.batch.to_numpy(b)
    assert isinstance(nb["roi"], np.ndarray) and isinstance(nb["image"][0], np.ndarray) and nb.meta is b.meta
    tb = dtr.batch.to_tensor(nb)
    assert torch.equal(tb["roi"], b["roi"].cpu()) and torch.equal(tb["image"][1], b["image"][1].cpu())
    assert isinstance(b["roi"], torch.Tensor)  # the input batch is untouched (shallow copy)


def _video_batch(n, seq):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

    rng = np.random.default_rng(31)
    raw = []
    for i in range(n):
        lab = cases.make_labels(rng, 450, 450)
        lab.pop("shapeparam")
        raw.append(dict(image=cases.make_image(rng, 450, 450, "smooth" if i % 2 else "noise"), **lab))
    cats = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")
    data = {"image": torch.from_numpy(np.stack([r["image"] for r in raw])).cuda()}
    for k in ("roi", "coord", "pose", "pt3d_68"):
        data[k] = torch.from_numpy(np.stack([r[k] for r in raw])).cuda()
    data["hasface"] = torch.tensor([bool(i % 3) for i in range(n)]).cuda()
    meta = Metadata((450, 450), 0, "video", list(seq), {k: FieldCategory(v) for k, v in cats.items()})
    return Batch(meta, data), raw, cats


def test_video_batch_one_draw_per_clip():
    """meta.seq: every frame of a clip gets the crop draw of the clip's first frame (geometric.py:180-191), through
    RandomFocusRoi and through FusedPoseAugmentation; pixels and labels against the oracle with those per-clip draws."""
    import trackertraincode_b200.datatransformation as dtr
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation

    seq = [0, 3, 5, 9]
    n = seq[-1]
    batch, raw, cats = _video_batch(n, seq)
    assert batch.meta.prefixshape == (n,)
    # ---- RandomFocusRoi on the clip batch
    torch.manual_seed(11)
    np.random.seed(11)
    focus = dtr.batch.RandomFocusRoi(S, rotation_aug_angle=30.0)
    seen = {}
    inner = focus.make_randomization_parameters
    focus.make_randomization_parameters = lambda B: seen.setdefault("p", inner(B))
    img_in = batch["image"]
    out = focus(batch)
    p = seen["p"]
    for a, b in zip(seq[:-1], seq[1:]):
        assert bool((p.scales[a:b] == p.scales[a]).all()) and bool((p.angles[a:b] == p.angles[a]).all())
        assert bool((p.translations[a:b] == p.translations[a]).all())
    assert len({float(s) for s in p.scales}) == 3  # three clips, three draws
    cs = torch.stack([torch.cos(p.angles), torch.sin(p.angles)], -1).numpy()
    for i in range(n):
        s = Sample((450, 450), {"image": raw[i]["image"][..., None], **{k: raw[i][k] for k in ("roi", "coord", "pose", "pt3d_68")}}, dict(cats))
        want, _ = ogeo.focus_roi(s, ogeo.RoiFocusParams(float(p.scales[i]), float(p.angles[i]), tuple(p.translations[i].tolist()),
                                                         cs_sn=(cs[i, 0], cs[i, 1])), S)
        assert np.array_equal(out["image"][i].cpu().numpy(), want.data["image"]), f"frame {i}"
        _close(out["pt3d_68"][i], want.data["pt3d_68"], 1e-4, 2e-4)
        _close(out["roi"][i], want.data["roi"], 1e-4, 2e-4)
    # ---- the fused class: per-clip crop AND flip draws, bool field smoothed
    batch, raw, cats = _video_batch(n, seq)
    aug = FusedPoseAugmentation(S, rotation_aug_angle=30.0, device="cuda", seed=5)
    torch.manual_seed(12)
    np.random.seed(12)
    d = aug._account_for_video(batch.meta, aug.draw(n))
    for a, b in zip(seq[:-1], seq[1:]):
        assert bool((d.geo.scales[a:b] == d.geo.scales[a]).all()) and bool((d.do_flip[a:b] == d.do_flip[a]).all())
        assert bool((d.rot_dir[a:b] == d.rot_dir[a]).all()) and bool((d.geo.angles[a:b] == d.geo.angles[a]).all())
    res = aug(batch, params=d)
    assert res["hasface"].dtype == torch.float32
    assert np.allclose(res["hasface"].cpu().numpy(), [0.9 if i % 3 else 0.1 for i in range(n)])  # normalization.py:26-30
    gp = opipe.GeoParams(d.geo.scales.numpy(), d.geo.angles.numpy(), d.geo.translations.numpy(), d.do_flip.numpy().astype(bool), d.rot_dir.numpy())
    ph = d.photo
    pp = opho.PhotoParams(list(ph.order), ph.apply.numpy(), ph.bits.numpy(), ph.gamma.numpy(), ph.contrast.numpy(), ph.brightness.numpy(),
                          ph.noise_apply.numpy(), ph.noise_std, ph.seed, ph.sample_offset, ph.clip)
    samples = [Sample((450, 450), {"image": r["image"][..., None], **{k: r[k] for k in ("roi", "coord", "pose", "pt3d_68")}}, dict(cats)) for r in raw]
    want, _ = opipe.augment_batch(samples, gp, pp, S)
    assert np.abs(res["image"].cpu().numpy() - want["image"]).max() <= 1.0 / 255
    for k in ("roi", "coord", "pt3d_68"):
        _close(res[k], want[k], 1e-4, 2e-5, k)
    _quat_close(res["pose"], want["pose"])


def test_noise_with_clipping_vs_oracle():
    """RandomGaussianNoiseWithClipping (intensity.py:43-53): the stage's output is clamped on the samples it touched, the
    others pass through un-clamped; a plain noise stage behind it is not clamped."""
    from trackertraincode_b200.datatransformation import batch as dtb

    rng = np.random.default_rng(2)
    n, H, W = 6, 33, 47
    x = (rng.random((n, 1, H, W)) * 1.6 - 0.3).astype(np.float32)  # values outside [0, 1] on purpose
    k = dtb.KorniaImageDistortions(dtb.RandomGaussianNoiseWithClipping(std=0.2, p=0.5), dtb.RandomGaussianNoise(std=0.05, p=0.5), seed=9)
    torch.manual_seed(4)
    p = k.draw(n)
    assert tuple(p.noise_clip[:2]) == (True, False) and not p.clip
    p.noise_apply[0] = torch.tensor([True, False, False, False])
    p.noise_apply[1] = torch.tensor([False, True, False, False])
    p.noise_apply[2] = torch.tensor([True, True, False, False])
    p.noise_apply[3] = False
    got = dtb.photometric_f32(torch.from_numpy(x).cuda(), p).cpu().numpy()
    pp = opho.PhotoParams([], p.apply.numpy(), p.bits.numpy(), p.gamma.numpy(), p.contrast.numpy(), p.brightness.numpy(), p.noise_apply.numpy(),
                          p.noise_std, p.seed, p.sample_offset, p.clip, tuple(p.noise_clip))
    want = opho.photometric_batch(x, pp)
    assert np.abs(got - want).max() <= 2e-4
    assert got[0].min() >= 0.0 and got[0].max() <= 1.0          # clipped stage applied
    assert got[1].min() < 0.0 and got[1].max() > 1.0            # only the plain stage: no clamp
    assert np.array_equal(got[3], x[3])                          # untouched sample passes through


def test_degenerate_roi_raises_deferred():
    """An empty view box zero-fills the crop; the mirror transforms report it (NativeError) at the next call / flush instead
    of training on it silently (the reference's cv2.resize raises on the spot)."""
    import trackertraincode_b200.datatransformation as dtr
    from trackertraincode_b200 import _native as N

    cs = [cases.make_case(i) for i in (0, 1, 5)]
    cs[1]["roi"] = np.float32([80, 80, 80, 80])
    focus = dtr.batch.FocusRoi(S, 1.1)
    with pytest.raises(N.NativeError, match="sample 1: empty view box"):
        focus(_frames(cs))    # (raises here already when the launch has finished by the time its status is looked at)
        focus.status.flush()
    ok = dtr.batch.FocusRoi(S, 1.1)
    ok(_frames([cases.make_case(0)]))
    ok.status.flush()


def test_prepared_call_is_bound_to_its_stream_unless_private():
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    cs = [cases.make_case(i) for i in (0, 1)]
    geo = E.GeoParams(torch.tensor([1.1, 1.2]), torch.tensor([0.3, 0.0]), torch.zeros(2, 2))
    shared = E.prepare_fused(_frames(cs), flags=N.F_FOCUS, out_size=S, geo=geo)
    private = E.prepare_fused(_frames(cs), flags=N.F_FOCUS, out_size=S, geo=geo, private_scratch=True)
    other = torch.cuda.Stream()
    with pytest.raises(N.NativeError, match="private_scratch"):
        shared.launch(other.cuda_stream)
    a = shared.launch().batch["image"]
    other.wait_stream(torch.cuda.current_stream())
    private.launch_plan(other.cuda_stream)      # the two phases on a second stream, as a pipelined loop runs them
    private.launch_main(other.cuda_stream)
    other.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(a, private.result.batch["image"])
