"""Host-side mirror of the reference API (no GPU): Batch / Metadata / Collation (datasets/batch.py:15-238), Affine2d
(neuralnets/affine2d.py) against vectors produced by the unmodified reference, the parameter samplers
(batch/geometric.py:58-96, pipelines.py:510-527), loader glue (datatransformation/loader.py) and launch scheduling."""
import os

import numpy as np
import pytest
import torch

from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
from trackertraincode_b200.datatransformation import (FusedPoseAugmentation, SampleBySampleLoader, SegmentedCollationDataLoader,
                                                     TransformedDataset, _engine as E, batch as dtb, sharding)
from trackertraincode_b200.datatransformation.fused import draw_photo_params
from trackertraincode_b200.neuralnets.affine2d import Affine2d

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CATS = {"image": FieldCategory.image, "roi": FieldCategory.roi, "pose": FieldCategory.quat}


def sample(i, wh=(20, 16), tag="a"):
    w, h = wh
    return Batch(Metadata((w, h), 0, tag, None, dict(CATS)),
                 {"image": torch.full((h, w, 1), i, dtype=torch.uint8), "roi": torch.tensor([1.0, 2, 3, 4]) + i,
                  "pose": torch.tensor([0.0, 0, 0, 1])})


def test_batch_dict_protocol_and_views():
    b = sample(3)
    assert b.meta.prefixshape == () and b.meta.is_single_frame and b.meta.image_wh == (20, 16)
    assert set(b.keys()) == {"image", "roi", "pose"} and "roi" in b and b.get_category("pose") == FieldCategory.quat
    bb = b.with_batchdim()
    assert bb.meta.batchsize == 1 and bb["roi"].shape == (1, 4) and b.meta.batchsize == 0  # the original is untouched
    frames = list(Batch.collate([sample(0), sample(1), sample(2)]).iter_frames())
    assert len(frames) == 3 and float(frames[2]["roi"][0]) == 3.0 and frames[0].meta.batchsize == 0
    import copy

    c = copy.copy(b)
    c["roi"] = b["roi"] + 1
    assert float(b["roi"][0]) == 4.0  # shallow copy: replacing an entry does not touch the source


def test_collation_stacked_segmented_and_ragged():
    out = Batch.collate([sample(i) for i in range(4)])
    assert out.meta.batchsize == 4 and out["image"].shape == (4, 16, 20, 1) and out["roi"].shape == (4, 4)
    seg = Batch.Collation(lambda b: b.meta.tag)([sample(0, tag="a"), sample(1, tag="b"), sample(2, tag="a")])
    assert isinstance(seg, list) and sorted(x.meta.batchsize for x in seg) == [1, 2]
    assert {x.meta.tag for x in seg} == {"a", "b"}
    # equal-size requirement of the stacked form ...
    with pytest.raises(RuntimeError):
        Batch.collate([sample(0), sample(1, wh=(24, 16))])
    # ... which the ragged form lifts: image fields travel as a list of per-sample tensors
    rag = Batch.Collation(ragged_images=True)([sample(0), sample(1, wh=(24, 16)), sample(2, wh=(8, 8))])
    assert rag.meta.batchsize == 3 and isinstance(rag["image"], list) and [tuple(t.shape) for t in rag["image"]] == [(16, 20, 1), (16, 24, 1), (8, 8, 1)]
    assert rag["roi"].shape == (3, 4)
    # collating collated ragged batches concatenates the lists
    rag2 = Batch.Collation(ragged_images=True)([rag, rag])
    assert rag2.meta.batchsize == 6 and len(rag2["image"]) == 6


def test_sequence_collation_offsets():
    def clip(n, tag="v"):
        return Batch(Metadata((8, 8), 0, tag, [0, n], dict(CATS)), {"roi": torch.zeros(n, 4), "pose": torch.zeros(n, 4),
                                                                    "image": torch.zeros(n, 8, 8, 1, dtype=torch.uint8)})

    out = Batch.collate([clip(3), clip(2), clip(4)])
    assert out.meta.seq == [0, 3, 5, 9] and out.meta.prefixshape == (9,) and out.meta.sequence_start_end == [(0, 3), (3, 5), (5, 9)]
    assert [s["roi"].shape[0] for s in out.iter_sequences()] == [3, 2, 4]


def test_affine2d_against_reference_vectors():
    """Affine2d (host mirror of neuralnets/affine2d.py) vs outputs of the reference's own class (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "algebra.npz"))
    n = len(g["angles"])
    t, ang, sc = (torch.from_numpy(g[k]) for k in ("translations", "angles", "scales_in"))
    for i in range(n):
        a = Affine2d.trs(translations=t[i], angles=ang[i], scales=sc[i])
        if i % 3 == 0:
            a = a @ Affine2d.range_remap_2d([0.0, 0.0], [129, 129], [129, 0], [0, 129])
        b = Affine2d.trs(translations=t[(i + 1) % n], angles=ang[(i + 5) % n], scales=sc[(i + 3) % n])
        np.testing.assert_allclose(a.tensor().numpy(), g["mats"][i], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose((a @ b).tensor().numpy(), g["prods"][i], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(a.inv().tensor().numpy(), g["invs"][i], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(a.scales.numpy(), g["scales"][i], rtol=1e-6)
        np.testing.assert_allclose(a.det.numpy(), g["dets"][i], rtol=1e-5, atol=1e-6)


def test_geometric_samplers_distribution():
    torch.manual_seed(0)
    np.random.seed(0)
    p = dtb.MakeRoiRandomizationParameters(30.0, 1.1)((20000,))
    assert p.scales.shape == (20000,) and p.translations.shape == (20000, 2) and (p.upfilter, p.downfilter) == ("linear", "area")
    assert abs(float(p.scales.mean()) - 1.1) < 5e-3 and float(p.scales.min()) >= 0.6 - 1e-6 and float(p.scales.max()) <= 1.6 + 1e-6
    assert float(p.translations.abs().max()) <= 1.0 and abs(float(p.translations.std()) - 0.47) < 0.03  # clipped 0.5 N(0,1)
    rot = p.angles != 0
    assert abs(float(rot.float().mean()) - 1 / 3) < 0.015
    assert torch.allclose(p.angles[rot].abs(), torch.tensor(np.pi / 6, dtype=p.angles.dtype))
    assert abs(float((p.angles[rot] > 0).float().mean()) - 0.5) < 0.03
    single = dtb.MakeRoiRandomizationParameters(30.0, 1.1)(())
    assert single.scales.shape == () and single.translations.shape == (2,)
    e = dtb.NoRoiRandomization(1.2)((5,))
    assert torch.all(e.scales == 1.2) and not e.angles.any() and not e.translations.any()


def test_photometric_sampler_matches_pipeline_probabilities():
    torch.manual_seed(1)
    p = draw_photo_params(20000, seed=7, sample_offset=123)
    assert len(p.order) == 4 and len(set(p.order)) == 4 and p.seed == 7 and p.sample_offset == 123
    chosen = torch.zeros(6, dtype=torch.bool)
    chosen[list(p.order)] = True
    want = torch.tensor([0.2, 0.01, 0.2, 0.2, 0.2, 0.1]) * chosen
    assert torch.allclose(p.apply.float().mean(0), want, atol=0.012)
    assert torch.allclose(p.noise_apply.float().mean(0), torch.tensor([0.25, 0.25**2, 0.25**3, 0.25**4]), atol=0.012)
    assert int(p.bits.min()) >= 4 and int(p.bits.max()) <= 5
    assert 0.5 <= float(p.gamma.min()) and float(p.gamma.max()) <= 2.0 and 0.7 <= float(p.contrast.min()) and float(p.brightness.max()) <= 1.5


def test_kornia_container_descriptor_rules():
    k = dtb.KorniaImageDistortions(dtb.RandomEqualize(p=0.2), dtb.RandomPosterize((4.0, 6.0), p=0.01), dtb.RandomGamma((0.5, 2.0), p=0.2),
                                   dtb.RandomContrast((0.7, 1.5), p=0.2), dtb.RandomBrightness((0.7, 1.5), p=0.2),
                                   dtb.RandomGaussianBlur(p=0.1, kernel_size=(5, 5), sigma=(1.5, 1.5), silence_instantiation_warning=True),
                                   random_apply=4)
    d = k.draw(64)
    assert len(d.order) == 4 and not d.clip and not d.noise_apply.any()
    assert not d.apply[:, [i for i in range(6) if i not in d.order]].any()
    k2 = dtb.KorniaImageDistortions(dtb.RandomGaussianNoise(std=4 / 255, p=1.0), dtb.RandomGaussianNoise(std=16 / 255, p=0.0), dtb.OnlyClip(p=1.0))
    d2 = k2.draw(8)
    assert d2.order == [] and d2.clip and d2.noise_apply[:, 0].all() and not d2.noise_apply[:, 1:].any()
    assert d2.noise_std[:2] == (4 / 255, 16 / 255)
    from trackertraincode_b200._native import NativeError

    with pytest.raises(NativeError):
        dtb.KorniaImageDistortions(dtb.RandomGaussianNoise(std=0.1), dtb.RandomGamma((0.5, 2.0)))  # point op after noise
    with pytest.raises(NativeError):
        dtb.RandomGaussianBlur(kernel_size=(3, 3), sigma=(1.0, 1.0))
    with pytest.raises(NativeError):
        dtb.PutRoiFromLandmarks(extend_to_forehead=True)
    with pytest.raises(NativeError):
        FusedPoseAugmentation(129, roi_override="extent_to_forehead")


def test_launch_order_first_wave_unrotated_then_by_cost():
    """Rotated samples get their canvas from the canvas workers, which start with the first wave of the fused kernel: the
    first wave holds the dearest unrotated samples, everything else follows by cost."""
    B = 8
    geo = E.GeoParams(torch.ones(B), torch.tensor([0, 0.5, 0, 0, 0, 0, -0.5, 0.0]), torch.zeros(B, 2))
    assert E.launch_order(B, E.GeoParams(torch.ones(B), torch.zeros(B), torch.zeros(B, 2)), None) is None
    o = E.launch_order(B, geo, None)
    assert sorted(o.tolist()) == list(range(B)) and set(o[-2:].tolist()) == {1, 6}  # (all 6 unrotated fit the first wave)
    ph = draw_photo_params(B, 0, 0)
    ph.order, ph.apply = [5, 0, 2, 3], torch.zeros(B, 6, dtype=torch.bool)
    ph.apply[3, 5] = True  # blurred: the dearest unrotated sample
    ph.apply[6, 5] = True  # rotated and blurred: first of the rotated ones
    ph.noise_apply = torch.zeros(B, 4, dtype=torch.bool)
    o = E.launch_order(B, geo, ph)
    assert o[0].item() == 3 and o[-2:].tolist() == [6, 1]
    old = E.FIRST_WAVE_SAMPLES
    try:
        E.FIRST_WAVE_SAMPLES = 2  # a first wave of two clusters: blurred + one plain, then the rotated blurred one
        o = E.launch_order(B, geo, ph)
        assert o[0].item() == 3 and o[2].item() == 6 and not {1, 6} & set(o[:2].tolist())
    finally:
        E.FIRST_WAVE_SAMPLES = old


class _DS(torch.utils.data.Dataset):
    def __len__(self):
        return 10

    def __getitem__(self, i):
        return sample(i, wh=(20 + 2 * (i % 3), 16), tag="even" if i % 2 == 0 else "odd")


def test_loader_glue_ragged_segmented():
    seen = []
    ds = TransformedDataset(_DS(), lambda b: b)
    ld = SegmentedCollationDataLoader(ds, batch_size=5, num_workers=0, segmentation_key_getter=lambda b: b.meta.tag, pin_memory=False,
                                      postprocess=lambda b: seen.append(b.meta.tag) or b, ragged_images=True)
    assert len(ld) == 2
    groups = list(ld)
    assert all(isinstance(g, list) for g in groups) and sum(b.meta.batchsize for g in groups for b in g) == 10
    assert set(seen) == {"even", "odd"}
    for b in ld.iter_unrolled():
        assert isinstance(b["image"], list) and len(b["image"]) == b.meta.batchsize and b["roi"].shape == (b.meta.batchsize, 4)
    items = list(SampleBySampleLoader(_DS(), num_workers=0, postprocess=lambda b: b))
    assert len(items) == 10 and items[3].meta.tag == "odd"


def test_loader_lookahead_keeps_order_and_runs_ahead():
    from trackertraincode_b200.datatransformation import PostprocessingLoader
    issued = []
    ld = PostprocessingLoader(list(range(12)), batch_size=3, postprocess=lambda t: issued.append(int(t[0])) or t, lookahead=2)
    assert len(ld) == 4 and ld.dataset == list(range(12))
    it = iter(ld)
    first = next(it)
    assert first.tolist() == [0, 1, 2] and issued == [0, 3, 6]  # two groups already went through the hook
    assert [t.tolist()[0] for t in it] == [3, 6, 9] and issued == [0, 3, 6, 9]
    plain = PostprocessingLoader(list(range(4)), batch_size=2)  # no hook: items pass through
    assert [t.tolist() for t in plain] == [[0, 1], [2, 3]]
    with pytest.raises(TypeError):
        list(SegmentedCollationDataLoader.__new__(SegmentedCollationDataLoader)._emit(3))


def test_sharding_helpers_single_process():
    for n, w in [(512, 8), (10, 3), (7, 8)]:
        r = [sharding.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    ids = {sharding.global_sample_offset(s, k, 4, 256) for s in range(3) for k in range(4)}
    assert ids == {256 * i for i in range(12)}
    assert sharding.max_over_ranks(1.5) == 1.5


def test_fused_augmentation_video_draws_and_row_bands():
    """Host side of FusedPoseAugmentation: clips share the draw of their first frame (geometric.py:180-191); the row bands
    handed to b200aug_upload_row_bands cover the rows of the oracle's integer view boxes (rotated samples: of the rotated
    square), with margin, for every sample."""
    from oracle import geometric as ogeo

    aug = FusedPoseAugmentation(129, rotation_aug_angle=30.0, device="cpu")
    torch.manual_seed(0)
    np.random.seed(0)
    B = 512
    d = aug.draw(B)
    meta = Metadata((450, 450), 0, "v", (0, 100, 101, 512), {})
    d = aug._account_for_video(meta, d)
    for a, b in meta.sequence_start_end:
        for t in (d.geo.scales, d.geo.angles, d.geo.translations, d.do_flip, d.rot_dir, d.geo.cos_sin):
            assert bool((t[a:b] == t[a:a + 1]).all())
    torch.manual_seed(1)
    np.random.seed(1)
    d = aug.draw(B)
    rng = np.random.default_rng(3)
    wh = rng.uniform(147, 250, (B, 2))
    c = 225 + rng.uniform(-140, 140, (B, 2))  # many boxes hang over the frame
    roi = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    boxes = aug._touched_boxes(torch.from_numpy(roi), d, 450, 450)
    lo, hi = boxes[:, 1], boxes[:, 3]
    v = ogeo.round_view_roi(ogeo.compute_view_roi(roi, d.geo.scales.numpy(), d.geo.translations.numpy())).astype(np.int64)
    rot = d.geo.angles.numpy() != 0
    side = (v[:, 2] - v[:, 0]).astype(np.float64)
    cy = 0.5 * (v[:, 1] + v[:, 3])
    ang = np.abs(d.geo.angles.numpy().astype(np.float64))
    half = np.where(rot, 0.5 * side * (np.cos(ang) + np.sin(ang)) + 1.0, 0.5 * (v[:, 3] - v[:, 1]))
    need_lo = np.clip(np.floor(cy - half), 0, 450)
    need_hi = np.clip(np.ceil(cy + half) + 1, 0, 450)  # + 1: the second bilinear tap row
    touched = need_hi > need_lo
    assert (lo[touched] <= need_lo[touched]).all() and (hi[touched] >= need_hi[touched]).all()
    assert lo.dtype == np.int32 and (hi >= lo).all() and (hi - lo).sum() < 0.75 * B * 450
    # the same along x (the boxes of b200aug_upload_boxes); columns come in multiples of 16
    cx = 0.5 * (v[:, 0] + v[:, 2])
    halfx = np.where(rot, 0.5 * side * (np.cos(ang) + np.sin(ang)) + 1.0, 0.5 * (v[:, 2] - v[:, 0]))
    need_x0, need_x1 = np.clip(np.floor(cx - halfx), 0, 450), np.clip(np.ceil(cx + halfx) + 1, 0, 450)
    touched = touched & (need_x1 > need_x0)
    assert (boxes[touched, 0] <= need_x0[touched]).all() and (boxes[touched, 2] >= need_x1[touched]).all()
    assert (boxes[:, 0] % 16 == 0).all() and (boxes[:, 2] >= boxes[:, 0]).all() and boxes[:, 2].max() <= 450
    area = ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])).sum()
    assert area < 0.5 * B * 450 * 450
