"""GPU parity of the stand-alone entry points and of the full-size workload (through the C ABI):
  * b200aug_photometric_f32  -- KorniaImageDistortions called on its own on float images (batch/intensity.py:30-40)
  * PutRoiFromLandmarks      -- batch/misc.py:9-31
  * ragged collation -> FusedPoseAugmentation, the loader-side drop-in (loader.py:48-51, pipelines.py:508-543)
  * BASELINE.json config 2 at full size (512 x 450x450 -> 129x129): size-independent properties + oracle spot checks
"""
import numpy as np
import pytest
import torch

import cases
from oracle import geometric as ogeo, normalization as onrm, photometric as opho, pipeline as opipe
from oracle.geometric import Sample

pytestmark = pytest.mark.gpu
S = 129
CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")


def _photo(E, pp):
    return E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma),
                         torch.from_numpy(pp.contrast), torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply),
                         pp.noise_std, pp.seed, pp.sample_offset, pp.clip)


@pytest.mark.parametrize("hw", [(129, 129), (224, 288), (17, 5)])
def test_photometric_f32_vs_oracle(hw):
    """Arbitrary float images (not k/256), every op alone, in pairs around the blur / equalize, plus noise and clip."""
    from trackertraincode_b200.datatransformation import _engine as E, batch as dtb

    H, W = hw
    rng = np.random.default_rng(H * 1000 + W)
    n = 10
    x = rng.random((n, 1, H, W)).astype(np.float32)
    x[1] = np.round(x[1] * 4) / 4  # few levels
    x[2] = 0.25  # constant: equalize step 0
    x[3] = x[3] * 1.4 - 0.2  # values outside [0, 1]: histc ignores them, the clamps act
    orders = [[op] for op in range(6)] + [[5, 0], [0, 5], [2, 5, 0, 3], [1, 4, 5, 2], [3, 0, 4, 2], []]
    for k, order in enumerate(orders):
        pp = opho.sample_photo_params(rng, n, seed=17 + k, sample_offset=1000 * k)
        pp.order = order
        pp.apply[:] = True
        pp.apply[n - 1] = False  # one untouched sample
        pp.apply[3, 2] = False  # no gamma on the sample with negative values (pow -> NaN, which then poisons blur / equalize)
        pp.noise_apply[:] = False
        if k % 2:
            pp.noise_apply[:, k % 4] = True
        pp.clip = bool(k % 3)
        want = opho.photometric_batch(x, pp)
        got = dtb.photometric_f32(torch.from_numpy(x).cuda(), _photo(E, pp)).cpu().numpy()
        assert not np.isnan(got).any() and not np.isnan(want).any()
        err = np.abs(got - want).reshape(n, -1)
        if 0 in order and order.index(0) > 0 and (2 in order[:order.index(0)]):
            # equalize is a step function: a 1-ulp difference of powf() in front of it moves a pixel across a histogram
            # bin edge now and then, which shifts that pixel by a few grey levels.  Bound how often and by how much.
            assert (err > 2e-4).mean() < 2e-3 and err.max() <= 12.0 / 255, f"order {order}: {(err > 2e-4).mean()} {err.max()}"
        else:
            assert err.max() <= 2e-4, f"order {order}: {err.max(1)}"
        assert np.median(err.max(1)) <= 3e-5
    # NaN semantics of torch.pow / torch.clamp on negative input are kept by the stand-alone kernel
    pp = opho.sample_photo_params(rng, n, seed=9)
    pp.order, pp.apply[:], pp.noise_apply[:] = [2], True, False
    with np.errstate(invalid="ignore"):
        want = opho.photometric_batch(x, pp)
    got = dtb.photometric_f32(torch.from_numpy(x).cuda(), _photo(E, pp)).cpu().numpy()
    assert np.isnan(want[3]).any() and np.array_equal(np.isnan(got), np.isnan(want))
    # bias fuses whiten_batch; in-place (in == out) is allowed without blur
    pp = opho.sample_photo_params(rng, n, seed=3)
    pp.order, pp.apply[:], pp.noise_apply[:] = [3, 4], True, False
    got = dtb.photometric_f32(torch.from_numpy(x).cuda(), _photo(E, pp), bias=-0.5).cpu().numpy()
    assert np.abs(got - (opho.photometric_batch(x, pp) - np.float32(0.5))).max() <= 1e-6


def test_kornia_container_call_and_whiten():
    """The reference's two containers + whiten_batch as separate calls == the oracle on the same draws."""
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import batch as dtb

    rng = np.random.default_rng(2)
    n = 16
    u8 = np.stack([cases.make_image(rng, S, S, "noise" if i % 2 else "smooth") for i in range(n)])[:, None]
    x = u8.astype(np.float32) / np.float32(256)
    b = Batch(Metadata(S, n, "t", None, {"image": FieldCategory.image, "roi": FieldCategory.roi}),
              {"image": torch.from_numpy(x).cuda(), "roi": torch.zeros(n, 4).cuda()})
    s1 = dtb.KorniaImageDistortions(dtb.RandomEqualize(p=0.5), dtb.RandomPosterize((4.0, 6.0), p=0.5), dtb.RandomGamma((0.5, 2.0), p=0.5),
                                    dtb.RandomContrast((0.7, 1.5), p=0.5), dtb.RandomBrightness((0.7, 1.5), p=0.5),
                                    dtb.RandomGaussianBlur(p=0.5, kernel_size=(5, 5), sigma=(1.5, 1.5)), random_apply=4)
    s2 = dtb.KorniaImageDistortions(dtb.RandomGaussianNoise(std=4 / 255, p=0.5), dtb.RandomGaussianNoise(std=16 / 255, p=0.5),
                                    dtb.RandomGaussianNoise(std=32 / 255, p=0.25), dtb.RandomGaussianNoise(std=64 / 255, p=0.25),
                                    dtb.OnlyClip(p=1.0), seed=5)
    torch.manual_seed(0)
    d1, d2 = s1.draw(n), s2.draw(n)
    out = dtb.whiten_batch(s2(s1(b, params=d1), params=d2))
    assert out["roi"] is b["roi"] and out.meta is b.meta  # non-image fields pass through, like the reference's shallow copy
    pp = opho.PhotoParams(list(d1.order), d1.apply.numpy(), d1.bits.numpy(), d1.gamma.numpy(), d1.contrast.numpy(), d1.brightness.numpy(),
                          d2.noise_apply.numpy(), d2.noise_std, d2.seed, d2.sample_offset, True)
    want = onrm.whiten_image(opho.photometric_batch(x, pp))
    err = np.abs(out["image"].cpu().numpy() - want).max()
    assert err <= 2e-4, err
    # unseeded calls draw fresh parameters and advance the noise stream
    a1 = s2(b)["image"]
    a2 = s2(b)["image"]
    assert s2.samples_seen == 3 * n and not torch.equal(a1, a2)


def test_put_roi_from_landmarks():
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import batch as dtb

    rng = np.random.default_rng(4)
    pts = (rng.random((7, 68, 3)) * 300).astype(np.float32)
    cats = {"pt3d_68": FieldCategory.points, "roi": FieldCategory.roi, "pose": FieldCategory.quat}
    b = Batch(Metadata((450, 450), 7, None, None, dict(cats)), {"pt3d_68": torch.from_numpy(pts).cuda(), "roi": torch.zeros(7, 4).cuda(),
                                                               "pose": torch.zeros(7, 4).cuda()})
    out = dtb.PutRoiFromLandmarks()(b)
    want = np.concatenate([pts[..., :2].min(1), pts[..., :2].max(1)], -1)
    assert np.array_equal(out["roi"].cpu().numpy(), want)
    single = Batch(Metadata((450, 450), 0, None, None, {"pt3d_68": FieldCategory.points}), {"pt3d_68": torch.from_numpy(pts[3]).cuda()})
    o = dtb.PutRoiFromLandmarks()(single)  # the reference creates the key when it is missing (misc.py:29-30)
    assert o["roi"].shape == (4,) and np.array_equal(o["roi"].cpu().numpy(), want[3])
    nolm = Batch(Metadata((450, 450), 0, None, None, {}), {"x": torch.zeros(3).cuda()})
    assert "roi" not in dtb.PutRoiFromLandmarks()(nolm)


def _raw_sample(rng, i, wh):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

    lab = cases.make_labels(rng, *wh)
    lab.pop("shapeparam")
    img = cases.make_image(rng, wh[0], wh[1], "noise" if i % 2 else "smooth")
    cats = {k: FieldCategory(v) for k, v in CATS.items()}
    data = {"image": torch.from_numpy(img[..., None].copy())}
    data.update({k: torch.from_numpy(v) for k, v in lab.items()})
    return Batch(Metadata(wh, 0, "w%d" % wh[0], None, cats), data), dict(image=img, **lab)


def test_ragged_collation_into_fused_augmentation():
    """Workers return raw frames of different sizes; the collated ragged batch goes through FusedPoseAugmentation as the
    loader's postprocess; replaying the draws through the oracle gives the same crops and labels."""
    from trackertraincode_b200.datasets.batch import Batch
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation

    rng = np.random.default_rng(8)
    sizes = [(450, 450), (320, 240), (640, 480), (200, 180)]
    pairs = [_raw_sample(rng, i, sizes[i % 4]) for i in range(12)]
    collated = Batch.Collation(ragged_images=True)([p[0] for p in pairs])
    assert isinstance(collated["image"], list)
    aug = FusedPoseAugmentation(S, rotation_aug_angle=30.0, device="cuda", seed=21)
    torch.manual_seed(3)
    np.random.seed(3)
    draws = aug.draw(12)
    out = aug(collated.pin_memory(), params=draws)
    assert out["image"].shape == (12, 1, S, S) and out["image"].dtype == torch.float32 and out.meta.imagesize == S
    gp = opipe.GeoParams(draws.geo.scales.numpy(), draws.geo.angles.numpy(), draws.geo.translations.numpy(),
                         draws.do_flip.numpy().astype(bool), draws.rot_dir.numpy())
    d = draws.photo
    pp = opho.PhotoParams(list(d.order), d.apply.numpy(), d.bits.numpy(), d.gamma.numpy(), d.contrast.numpy(), d.brightness.numpy(),
                          d.noise_apply.numpy(), d.noise_std, d.seed, d.sample_offset, d.clip)
    samples = [Sample(sizes[i % 4], {k: (v[..., None] if k == "image" else v) for k, v in raw.items()}, CATS) for i, (_, raw) in enumerate(pairs)]
    want, _ = opipe.augment_batch(samples, gp, pp, S)
    err = np.abs(out["image"].cpu().numpy() - want["image"]).max()
    assert err <= 1.0 / 255, err
    for k in ("roi", "coord", "pt3d_68"):
        np.testing.assert_allclose(out[k].cpu().numpy(), want[k], rtol=1e-4, atol=2e-5, err_msg=k)
    q, qw = out["pose"].cpu().numpy(), want["pose"]
    assert np.minimum(np.abs(q - qw).max(-1), np.abs(q + qw).max(-1)).max() < 2e-5


def test_full_size_config2_properties():
    """B=512, 450x450 -> 129x129 (BASELINE.json configs[1]): size-independent properties, and the oracle's full chain on every
    one of the 512 samples."""
    import bench
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import _engine as E

    B = bench.BATCH
    host = bench.make_host_batch(5)
    gp, pp = bench.draw_params(55, B, 4096)
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    dev = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
    flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN
    an = torch.from_numpy(gp.angles)

    def run(lo, hi, **kw):
        sl = slice(lo, hi)
        b = Batch(Metadata((bench.SRC, bench.SRC), hi - lo, "t", None, dict(cats)), {k: v[sl] for k, v in dev.items()})
        p = pp.slice(lo, hi)
        geo = E.GeoParams(torch.from_numpy(gp.scales[sl]), an[sl], torch.from_numpy(gp.translations[sl]), E.host_cos_sin(an[sl]))
        r = E.fused_forward(b, flags=kw.pop("flags", flags), out_size=S, geo=geo, do_flip=torch.from_numpy(gp.do_flip[sl].astype(np.uint8)),
                            rot_dir=torch.from_numpy(gp.rot_dir[sl]), photo=_photo(E, p), want_status=True, want_view_roi=True, **kw)
        assert not r.status.cpu().numpy().any()
        return r

    whole = run(0, B)
    img = whole.batch["image"]
    assert img.shape == (B, 1, S, S) and float(img.min()) >= -0.5 and float(img.max()) <= 0.5
    # (1) integers: the view boxes are bit-exact against the oracle for all 512 samples
    v = ogeo.round_view_roi(ogeo.compute_view_roi(host["roi"], gp.scales, gp.translations))
    assert np.array_equal(whole.view_roi.cpu().numpy(), v)
    # (2) determinism + independence of the launch geometry: repeat, other cluster sizes, no scheduling, no scratch canvas
    for kw in (dict(), dict(cluster_size=1), dict(cluster_size=4), dict(schedule=False), dict(use_workspace=False)):
        assert torch.equal(run(0, B, **kw).batch["image"], img), kw
    # (3) sharding invariance (what N GPUs do): the batch in 4 slices == the batch at once, incl. the Philox noise
    parts = [run(lo, lo + B // 4) for lo in range(0, B, B // 4)]
    assert torch.equal(torch.cat([p.batch["image"] for p in parts]), img)
    for k in ("roi", "coord", "pose", "pt3d_68"):
        assert torch.equal(torch.cat([p.batch[k] for p in parts]), whole.batch[k]), k
    # (4) flip is an involution on the geometric output: flipping the flipped crop gives the unflipped crop
    geo_flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT
    a = run(0, B, flags=geo_flags).batch["image"]
    saved = gp.do_flip.copy()
    gp.do_flip[:] = ~saved
    b = run(0, B, flags=geo_flags).batch["image"]
    gp.do_flip[:] = saved
    norot = torch.from_numpy(gp.rot_dir == 0).cuda()
    assert torch.equal(a[norot].flip(-1), b[norot])
    # (5) the oracle on all 512 samples, full chain
    idx = np.arange(B)
    img_host, pts_host = img.cpu().numpy(), whole.batch["pt3d_68"].cpu().numpy()
    worst = 0.0
    samples = [Sample((bench.SRC, bench.SRC), {k: (host[k][i][..., None] if k == "image" else host[k][i]) for k in bench.CATS}, bench.CATS) for i in idx]
    g = opipe.GeoParams(gp.scales[idx], gp.angles[idx], gp.translations[idx], gp.do_flip[idx], gp.rot_dir[idx])
    for j, i in enumerate(idx):
        p1 = pp.slice(int(i), int(i) + 1)
        g1 = opipe.GeoParams(*(x[j:j + 1] for x in (g.scales, g.angles, g.translations, g.do_flip, g.rot_dir)))
        want, _ = opipe.augment_batch(samples[j:j + 1], g1, p1, S)
        err = np.abs(img_host[i] - want["image"][0]).max()
        worst = max(worst, float(err))
        assert err <= 1.0 / 255, (i, err)
        np.testing.assert_allclose(pts_host[i], want["pt3d_68"][0], rtol=1e-4, atol=2e-5)
    assert worst <= 1e-5  # (observed 4e-6: the float32 photometric chain -- pow, Box-Muller -- agrees to a few ulp; 1/255 is the contract)


def test_upload_modes_agree():
    """Four ways of getting pinned host frames to the kernel give bit-identical results: whole frames copied
    (Batch.to), only the boxes the sampled view boxes can touch (b200aug_upload_boxes, 2-D copies) or their whole row bands
    (b200aug_upload_row_bands) -- the rest of the device frame stack holds stale pixels of the previous batches --, and frames
    read in place from pinned host memory.  Pageable host frames are refused by the engine."""
    import bench
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation, _engine as E

    B = 128
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    augs = {m: FusedPoseAugmentation(S, rotation_aug_angle=30.0, device="cuda", seed=5, zero_copy_frames=(m == "zero_copy"),
                                     upload_row_bands=(m in ("bands", "boxes"))) for m in ("copy", "bands", "boxes", "zero_copy")}
    augs["bands"].upload_boxes, augs["boxes"].upload_boxes = False, True
    for rnd in range(3):  # the band uploads of later rounds land in a frame stack full of the earlier rounds' rows
        host = bench.make_host_batch(9 + rnd, B)
        if rnd == 1:  # boxes hanging over the frame borders
            host["roi"][:, [1, 3]] += np.where(np.arange(B) % 2 == 0, -120.0, 150.0).astype(np.float32)[:, None]
            host["roi"][:, [0, 2]] += np.where(np.arange(B) % 3 == 0, -130.0, 140.0).astype(np.float32)[:, None]
        pinned = Batch(Metadata((bench.SRC, bench.SRC), B, "t", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in host.items()})
        outs = {}
        for m, aug in augs.items():
            torch.manual_seed(11 + rnd)
            np.random.seed(11 + rnd)
            draws = aug.draw(B)
            outs[m] = aug(pinned, params=draws)
        assert 0 < augs["bands"].uploaded_rows < 0.8 * B * bench.SRC
        assert 0 < augs["boxes"].uploaded_bytes < 0.8 * augs["bands"].uploaded_bytes
        for m in ("bands", "boxes", "zero_copy"):
            assert outs[m]["image"].is_cuda and torch.equal(outs[m]["image"], outs["copy"]["image"]), (rnd, m)
            for k in ("roi", "coord", "pose", "pt3d_68"):
                assert torch.equal(outs[m][k], outs["copy"][k]), (rnd, m, k)
    pageable = Batch(pinned.meta, {k: (torch.from_numpy(v) if k == "image" else torch.from_numpy(v).cuda()) for k, v in host.items()})
    with pytest.raises(N.NativeError):
        E.fused_forward(pageable, flags=N.F_NORMALIZE, out_size=S)


def test_randomised_sweep_script():
    """scripts/gpu_stress.py (ragged batches of odd frame sizes and pitches, rotations up to 45 degrees, output sizes 64 / 129 /
    200, oracle with identical draws): four rounds here, `profiles/r01d_stress.log` holds a 16-round run."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "gpu_stress.py"), "4"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
