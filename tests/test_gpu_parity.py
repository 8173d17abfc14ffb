"""GPU parity: the CUDA path (through the C ABI, via the Python mirror of the reference API) against
  (1) the golden vectors produced by the unmodified reference, and
  (2) the numpy/cv2 oracle on seeded random inputs.
Bars (BASELINE.json north_star): integer work bit-exact; pixels within 1/255 (we assert bit-exact on the uint8 crop);
landmarks / rotations within 1e-4 relative.
"""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import geometric as ogeo, normalization as onrm, photometric as opho, pipeline as opipe
from oracle.geometric import Sample

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts", shapeparam="")
LABELS = ("roi", "coord", "pose", "pt3d_68", "shapeparam")
S = 129


@pytest.fixture(scope="module")
def dtr():
    import trackertraincode_b200.datatransformation as dtr

    return dtr


@pytest.fixture(scope="module")
def chain():
    return np.load(os.path.join(GOLD, "focus_chain.npz"))


def make_batch(case_list, device="cuda"):
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

    cats = {k: FieldCategory(v) for k, v in CATS.items()}
    data = {"image": [torch.from_numpy(c["image"][..., None].copy()).to(device) for c in case_list]}
    for k in LABELS:
        data[k] = torch.from_numpy(np.stack([c[k] for c in case_list])).to(device)
    wh = case_list[0]["wh"]
    return Batch(Metadata(wh, len(case_list), "t", None, categories=cats), data)


def to_sample(c) -> Sample:
    data = {"image": c["image"][..., None]}
    for k in LABELS:
        data[k] = c[k]
    return Sample(c["wh"], data, dict(CATS))


def host_trig(angles):
    a = torch.as_tensor(np.asarray(angles, np.float32))
    return torch.stack([torch.cos(a), torch.sin(a)], -1)


def assert_labels(got, want, k, rtol=1e-4, atol=2e-5):
    g, w = got.cpu().numpy(), np.asarray(want)
    if k == "pose":
        d = np.minimum(np.abs(g - w).max(-1), np.abs(g + w).max(-1))
        assert d.max() < 2e-5, f"{k}: {d.max()}"
    else:
        np.testing.assert_allclose(g, w, rtol=rtol, atol=atol, err_msg=k)


def geo_from_cases(cs, E):
    sc = torch.tensor([float(c["scale"]) for c in cs])
    an = torch.tensor([float(c["angle"]) for c in cs])
    tr = torch.from_numpy(np.stack([c["translation"] for c in cs]))
    return E.GeoParams(sc, an, tr, host_trig(an.numpy()))


def test_golden_stagewise(dtr, chain):
    """Focus -> flip/rot90 -> normalize as three separate launches against the reference's own outputs."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    cs = [cases.make_case(i) for i in range(cases.N_CASES)]
    b = dtr.batch.offset_points_by_half_pixel(make_batch(cs))
    r = E.fused_forward(b, flags=N.F_FOCUS, out_size=S, geo=geo_from_cases(cs, E), want_view_roi=True, want_status=True)
    assert np.array_equal(r.view_roi.cpu().numpy(), chain["view_roi"])  # integer work: bit exact
    assert not r.status.cpu().numpy().any()
    assert np.array_equal(r.tr.cpu().numpy(), chain["tr"]), "focus transform must be reproduced bit for bit"
    img = r.batch["image"].cpu().numpy()[:, 0]
    bad = [(i, int((img[i] != chain["focus_image"][i]).sum())) for i in range(len(cs)) if not np.array_equal(img[i], chain["focus_image"][i])]
    assert not bad, f"crop pixels differ from the reference in cases {bad}"
    for k in LABELS:
        assert_labels(r.batch[k], chain["focus_" + k], k, atol=2e-4)
    b2 = r.batch
    b2.meta._imagesize = S
    flips = (torch.tensor([c["do_flip"] for c in cs], dtype=torch.uint8), torch.tensor([c["rot_dir"] for c in cs], dtype=torch.int8))
    b3 = dtr.batch.horizontal_flip_and_rot_90(0.01, b2, draws=flips)
    assert np.array_equal(b3["image"].cpu().numpy()[:, 0], chain["flip_image"])
    for k in LABELS:
        assert_labels(b3[k], chain["flip_" + k], k, atol=2e-4)
    b4 = dtr.batch.normalize_batch(b3)
    assert b4["image"].dtype == torch.float32
    assert np.array_equal(b4["image"].cpu().numpy()[:, 0], chain["flip_image"].astype(np.float32) / 256)
    for k in LABELS:
        assert_labels(b4[k], chain["final_" + k], k)
    w = dtr.batch.whiten_batch(b4)
    assert np.array_equal(w["image"].cpu().numpy()[:, 0], chain["flip_image"].astype(np.float32) / 256 - np.float32(0.5))


def test_golden_fused_single_launch(dtr, chain):
    """The same chain as ONE launch (what the training loop runs) must give the same bits."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    cs = [cases.make_case(i) for i in range(cases.N_CASES)]
    flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_WHITEN
    r = E.fused_forward(make_batch(cs), flags=flags, out_size=S, geo=geo_from_cases(cs, E),
                        do_flip=torch.tensor([c["do_flip"] for c in cs], dtype=torch.uint8),
                        rot_dir=torch.tensor([c["rot_dir"] for c in cs], dtype=torch.int8), want_view_roi=True)
    assert np.array_equal(r.view_roi.cpu().numpy(), chain["view_roi"])
    want = chain["flip_image"].astype(np.float32) / 256 - np.float32(0.5)
    assert np.array_equal(r.batch["image"].cpu().numpy()[:, 0], want)
    for k in LABELS:
        assert_labels(r.batch[k], chain["final_" + k], k)


def test_device_trig_production_angles(dtr, chain):
    """Without host cos/sin the kernel's correctly rounded values must still be bit-exact for 0 and +-30 degrees."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    prod = (0.0, float(np.float32(np.pi * 30.0 / 180.0)))
    idx = [i for i in range(cases.N_CASES) if abs(float(cases.make_case(i)["angle"])) in prod]
    cs = [cases.make_case(i) for i in idx]
    g = geo_from_cases(cs, E)
    g.cos_sin = None
    b = dtr.batch.offset_points_by_half_pixel(make_batch(cs))
    r = E.fused_forward(b, flags=N.F_FOCUS, out_size=S, geo=g)
    assert np.array_equal(r.tr.cpu().numpy(), chain["tr"][idx])
    assert np.array_equal(r.batch["image"].cpu().numpy()[:, 0], chain["focus_image"][idx])


def random_inputs(n, seed, wh=(450, 450)):
    rng = np.random.default_rng(seed)
    cs = []
    for i in range(n):
        lab = cases.make_labels(rng, *wh)
        img = cases.make_image(rng, wh[0], wh[1], "noise" if i % 2 else "smooth")
        cs.append(dict(wh=wh, image=img, **lab))
    return cs, rng


@pytest.mark.parametrize("roi_mode", ["original", "landmarks"])
def test_full_chain_vs_oracle(dtr, roi_mode):
    """Config-2-shaped random batch, all stages incl. photometric, against the oracle with identical draws."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    n = 96
    cs, rng = random_inputs(n, 11)
    gp = opipe.sample_geo_params(rng, n)
    gp.rot_dir[:8] = [1, -1, 1, -1, 0, 0, 1, -1]
    pp = opho.sample_photo_params(rng, n, seed=0xC0FFEE1234, sample_offset=77)
    # make every op / combination appear: force masks on the first samples
    pp.order = [5, 0, 2, 3]  # blur, equalize, gamma, contrast  (equalize after blur on sample 0)
    pp.apply[:] = rng.random((n, 6)) < 0.35
    pp.apply[:, [1, 4]] = False
    pp.apply[0] = [1, 0, 1, 1, 0, 1]
    pp.noise_apply[:6] = [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 1, 1], [0, 0, 0, 0]]
    cs_sn = host_trig(gp.angles).numpy()

    # oracle (cv2 for pixels), feeding it the host-evaluated cos/sin like the reference would compute them
    want_imgs, want = [], {k: [] for k in LABELS}
    samples = [to_sample(c) for c in cs]
    outs = []
    for i, s in enumerate(samples):
        s = onrm.offset_points_by_half_pixel(s)
        if roi_mode == "landmarks":
            s = ogeo.put_roi_from_landmarks(s)
        s, _ = ogeo.focus_roi(s, ogeo.RoiFocusParams(gp.scales[i], gp.angles[i], gp.translations[i], tuple(cs_sn[i])), S)
        if roi_mode == "landmarks":
            s = ogeo.put_roi_from_landmarks(s)
        s = ogeo.horizontal_flip_and_rot_90(s, bool(gp.do_flip[i]), int(gp.rot_dir[i]))
        outs.append(onrm.normalize_sample(s))
    pre = np.stack([o.data["image"] for o in outs])  # [n,1,S,S] float32 = u8/256
    want_img = onrm.whiten_image(opho.photometric_batch(pre, pp))

    flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN
    if roi_mode == "landmarks":
        flags |= N.F_ROI_FROM_LANDMARKS
    photo = E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma),
                          torch.from_numpy(pp.contrast), torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply),
                          pp.noise_std, pp.seed, pp.sample_offset, pp.clip)
    geo = E.GeoParams(torch.from_numpy(gp.scales), torch.from_numpy(gp.angles), torch.from_numpy(gp.translations), torch.from_numpy(cs_sn))
    r = E.fused_forward(make_batch(cs), flags=flags, out_size=S, geo=geo, do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)),
                        rot_dir=torch.from_numpy(gp.rot_dir), photo=photo, want_status=True)
    assert not r.status.cpu().numpy().any()
    got = r.batch["image"].cpu().numpy()
    # geometric-only launch: uint8 crop bit-exact against cv2
    r2 = E.fused_forward(make_batch(cs), flags=flags & ~(N.F_PHOTOMETRIC | N.F_WHITEN), out_size=S, geo=geo,
                         do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)), rot_dir=torch.from_numpy(gp.rot_dir))
    assert np.array_equal(r2.batch["image"].cpu().numpy(), pre), "uint8 crop (x 1/256) differs from cv2"
    err = np.abs(got - want_img).reshape(n, -1).max(1)
    assert err.max() <= 1.0 / 255, f"photometric chain off by {err.max()} (samples {np.nonzero(err > 1/255)[0][:10]})"
    assert np.median(err) < 1e-6
    for k in LABELS:
        assert_labels(r.batch[k], np.stack([o.data[k] for o in outs]), k)


def test_photometric_each_op_vs_oracle(dtr):
    """One op at a time on a plain (already cropped) batch: isolates every stage-1 op and each noise stage."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import _engine as E

    rng = np.random.default_rng(5)
    n = 12
    u8 = np.stack([cases.make_image(rng, S, S, "noise" if i % 2 else "smooth") for i in range(n)])
    u8[2] = (u8[2] // 64) * 64  # few grey levels: equalize with a tiny step
    u8[3] = 17  # constant image: equalize step == 0
    for op in range(6):
        for order in ([op], [5, op] if op != 5 else [0, 5, 2]):
            pp = opho.sample_photo_params(rng, n, seed=op + 1)
            pp.order = order
            pp.apply[:] = True
            pp.noise_apply[:] = False
            pp.noise_apply[:, op % 4] = True
            want = opho.photometric_batch(u8[:, None].astype(np.float32) / np.float32(256), pp)
            b = Batch(Metadata(S, n, None, None, {"image": FieldCategory.image}), {"image": torch.from_numpy(u8[:, None]).cuda()})
            photo = E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma),
                                  torch.from_numpy(pp.contrast), torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply),
                                  pp.noise_std, pp.seed, pp.sample_offset, pp.clip)
            got = E.fused_forward(b, flags=N.F_NORMALIZE | N.F_PHOTOMETRIC, out_size=S, photo=photo).batch["image"].cpu().numpy()
            err = np.abs(got - want).reshape(n, -1).max(1)
            # point ops agree to float32 rounding; the noise stage uses the hardware log/sqrt/sin/cos approximations
            # (<= ~1e-4 of the [0,1] range at std 64/255, far inside the 1/255 bar)
            assert err.max() <= 2e-4, f"op {op} order {order}: max err {err} "
            assert np.median(err) <= 3e-5


def test_edge_cases(dtr):
    """Empty / degenerate boxes are reported per sample, big boxes beyond the border zero-fill, B=0 is a no-op."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    cs, rng = random_inputs(6, 3, wh=(200, 180))
    cs[0]["roi"] = np.float32([50, 50, 50, 50])  # empty
    cs[1]["roi"] = np.float32([-400, -300, -250, -150])  # entirely outside: all zeros
    cs[2]["roi"] = np.float32([10, 10, 139, 139])  # 129 x 129 at scale 1: copy path
    cs[3]["roi"] = np.float32([20, 20, 84.5, 84.5])  # -> 64.5: exact 2x up-scale region
    sc = np.float32([1.0, 1.0, 1.0, 1.0, 1.1, 3.9])
    geo = E.GeoParams(torch.from_numpy(sc), torch.zeros(6), torch.zeros(6, 2))
    r = E.fused_forward(make_batch(cs), flags=N.F_FOCUS, out_size=S, geo=geo, want_status=True, want_view_roi=True)
    st = r.status.cpu().numpy()
    img = r.batch["image"].cpu().numpy()[:, 0]
    assert st[0] == N.S_EMPTY_BOX and not img[0].any()
    assert st[1] == 0 and not img[1].any()
    for i in range(1, 6):
        assert st[i] == 0
        view = ogeo.round_view_roi(ogeo.compute_view_roi(cs[i]["roi"], sc[i], np.zeros(2, np.float32)))
        assert np.array_equal(r.view_roi.cpu().numpy()[i], view)
        assert np.array_equal(img[i], ogeo.croprescale_image(cs[i]["image"], view, (S, S))), f"sample {i}"
    # tiny row buffer -> the kernel falls back to its per-pixel path: same bits, no error
    r2 = E.fused_forward(make_batch(cs), flags=N.F_FOCUS, out_size=S, geo=geo, want_status=True, rowbuf_capacity=64)
    assert not r2.status.cpu().numpy()[1:].any()
    assert np.array_equal(r2.batch["image"].cpu().numpy(), r.batch["image"].cpu().numpy())


def test_nonsquare_output_localizer_shape(dtr):
    """Config 5 shape: 640x480 frames -> 288x224 crops (LocalizerNet input), primitives a7/a12 only."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    cs, rng = random_inputs(5, 21, wh=(640, 480))
    geo = E.GeoParams(torch.full((5,), 1.3), torch.zeros(5), torch.zeros(5, 2))
    r = E.fused_forward(make_batch(cs), flags=N.F_FOCUS | N.F_HALF_PIXEL, out_size=(288, 224), geo=geo, want_view_roi=True, want_status=True)
    assert not r.status.cpu().numpy().any()
    img = r.batch["image"].cpu().numpy()[:, 0]
    for i, c in enumerate(cs):
        s, inter = ogeo.focus_roi(onrm.offset_points_by_half_pixel(to_sample(c)), ogeo.RoiFocusParams(np.float32(1.3), 0.0, (0.0, 0.0)), (288, 224))
        assert np.array_equal(r.view_roi.cpu().numpy()[i], inter["view_roi"])
        assert np.array_equal(img[i], s.data["image"][0])
        assert_labels(r.batch["roi"][i], s.data["roi"], "roi", atol=2e-4)
        assert_labels(r.batch["pt3d_68"][i], s.data["pt3d_68"], "pt3d_68", atol=2e-4)


def test_inter_area_with_an_upscaling_axis(dtr):
    """The reference asks cv2.resize for INTER_AREA whenever the mean scale is < 1 (image_geometric_cv2.py:65-82); when one
    axis up-scales cv2 then runs its 2-tap kernel with the area-mode coefficient rule on both axes.  Non-square outputs
    (view side between the output height and width: the localizer's 288 x 224) hit this; square outputs cannot, the rounded
    view box is at most one pixel off square."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    sides = [262.0, 270.0, 275.0, 281.0, 286.0, 225.0, 240.0]
    cs, rng = random_inputs(len(sides), 33, wh=(640, 480))
    for c, side in zip(cs, sides):
        cx, cy = rng.uniform(250, 390), rng.uniform(200, 280)
        c["roi"] = np.array([cx - 100, cy - 100, cx + 100, cy + 100], np.float32)
    scales = torch.tensor([s / 200.0 for s in sides])
    geo = E.GeoParams(scales, torch.zeros(len(sides)), torch.zeros(len(sides), 2))
    r = E.fused_forward(make_batch(cs), flags=N.F_FOCUS, out_size=(288, 224), geo=geo, want_view_roi=True, want_status=True)
    assert not r.status.cpu().numpy().any()
    img = r.batch["image"].cpu().numpy()[:, 0]
    mixed = 0
    for i, c in enumerate(cs):
        s, inter = ogeo.focus_roi(to_sample(c), ogeo.RoiFocusParams(np.float32(scales[i].item()), 0.0, (0.0, 0.0)), (288, 224))
        v = inter["view_roi"]
        mixed += (v[2] - v[0] < 288) and (0.5 * (288 / (v[2] - v[0]) + 224 / (v[3] - v[1])) < 1.0)
        assert np.array_equal(r.view_roi.cpu().numpy()[i], v)
        assert np.array_equal(img[i], s.data["image"][0]), f"sample {i} (view box {v})"
    assert mixed >= 4


def test_integer_factor_and_frame_borders(dtr):
    """Exact 2x / 3x INTER_AREA (cv2's integer box path), crops crossing every frame border, with and without rotation."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    rois = [
        [10, 20, 268, 278],      # 258 -> 129: (a+b+c+d+2)>>2
        [30, 40, 417, 427],      # 387 -> 129: rint(sum * 1/9)
        [10, 20, 268, 407],      # 258 x 387: mixed integer factors
        [-40, 100, 200, 340],    # left border
        [300, 100, 540, 340],    # right border
        [100, -60, 340, 180],    # top border
        [100, 300, 340, 540],    # bottom border (touches the frame's last row)
        [-30, -30, 480, 480],    # larger than the frame on every side
        [200, 191, 449, 440],    # right edge exactly at the frame edge, near the last row
        [0, 0, 450, 450],        # the whole frame
    ]
    n = len(rois)
    cs, rng = random_inputs(n, 31)
    for c, r in zip(cs, rois):
        c["roi"] = np.float32(r)
    for angle in (0.0, float(np.float32(np.pi / 6)), float(np.float32(-0.2))):
        an = np.full(n, angle, np.float32)
        geo = E.GeoParams(torch.ones(n), torch.from_numpy(an), torch.zeros(n, 2), host_trig(an))
        r = E.fused_forward(make_batch(cs), flags=N.F_FOCUS, out_size=S, geo=geo, want_status=True, want_view_roi=True)
        assert not r.status.cpu().numpy().any()
        img = r.batch["image"].cpu().numpy()[:, 0]
        trig = host_trig(an).numpy()
        for i, c in enumerate(cs):
            s, inter = ogeo.focus_roi(to_sample(c), ogeo.RoiFocusParams(np.float32(1.0), an[i], np.zeros(2, np.float32), tuple(trig[i])), S)
            assert np.array_equal(r.view_roi.cpu().numpy()[i], inter["view_roi"])
            assert np.array_equal(img[i], s.data["image"][0]), f"angle {angle} roi {rois[i]}: {(img[i] != s.data['image'][0]).sum()} pixels differ"


def test_rotated_paths_agree(dtr):
    """The rotated samples' two implementations (canvas in the L2 scratch workspace vs. canvas rows produced one at a time)
    and the per-pixel fallback (tiny row buffer) must give identical bits."""
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datatransformation import _engine as E

    n = 48
    cs, rng = random_inputs(n, 77)
    gp = opipe.sample_geo_params(rng, n)
    gp.angles[::2] = np.float32(np.pi / 6) * np.where(np.arange(n)[::2] % 4 == 0, 1, -1)
    gp.angles[1] = np.float32(0.7)  # more than the training range
    an = torch.from_numpy(gp.angles)
    geo = E.GeoParams(torch.from_numpy(gp.scales), an, torch.from_numpy(gp.translations), E.host_cos_sin(an))
    outs = []
    for kw in (dict(use_workspace=True), dict(use_workspace=False), dict(use_workspace=True, rowbuf_capacity=64)):
        r = E.fused_forward(make_batch(cs), flags=N.F_FOCUS, out_size=S, geo=geo, want_status=True, **kw)
        assert not r.status.cpu().numpy().any()
        outs.append(r.batch["image"].cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[0], outs[2])
    for i in range(0, n, 6):
        trig = host_trig(gp.angles[i:i + 1]).numpy()[0]
        s, _ = ogeo.focus_roi(to_sample(cs[i]), ogeo.RoiFocusParams(gp.scales[i], gp.angles[i], gp.translations[i], tuple(trig)), S)
        assert np.array_equal(outs[0][i, 0], s.data["image"][0])
