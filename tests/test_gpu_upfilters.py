"""The cubic / Lanczos up-filters (image_geometric_cv2.py:65-82, 105-119) on the GPU, through the mirror API, against outputs
of the unmodified reference (tests/golden/upfilter.npz) and against the oracle's models of OpenCV's kernels.

Kernel == model bit for bit everywhere.  Against the reference: bit-exact too, except where the reference's wheel lets Intel
IPP replace OpenCV's cv2.resize(INTER_CUBIC) kernel -- there within one grey level (tests/test_oracle_upfilters.py)."""
import os

import numpy as np
import pytest
import torch

import cases
import upfilter_cases as uc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "upfilter.npz"))


def check(out, ref, model, ipp_cubic_resize, what):
    assert out.shape == ref.shape, what
    assert np.array_equal(out, model), f"{what}: kernel vs oracle model"
    d = np.abs(out.astype(int) - ref.astype(int))
    if ipp_cubic_resize:
        assert d.max() <= 1 and (d > 0).mean() < 0.08, f"{what}: kernel vs reference (IPP cubic)"
    else:
        assert d.max() == 0, f"{what}: kernel vs reference"


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_focus_with_upfilter(gold, kind):
    from oracle import geometric as geo, normalization as nrm
    from test_gpu_prefilter import focus_batch
    from test_oracle_golden import host_cos_sin, to_sample
    from trackertraincode_b200.datatransformation import batch as dtb

    by_size = {}
    for j, i in enumerate(uc.FOCUS_CASES):
        by_size.setdefault(cases.make_case(i)["wh"], []).append((j, i))
    for wh, items in by_size.items():
        cs, b = focus_batch([i for _, i in items])
        b = dtb.offset_points_by_half_pixel(b)
        params = dtb.RoiFocusRandomizationParameters(
            scales=torch.tensor([float(c["scale"]) for c in cs]), angles=torch.tensor([float(c["angle"]) for c in cs]),
            translations=torch.from_numpy(np.stack([c["translation"] for c in cs])), upfilter=kind, downfilter="area")
        aug = dtb.RandomFocusRoi(cases.OUT_SIZE)
        aug.make_randomization_parameters = lambda B, params=params: params
        res = aug(b)
        aug.status.flush()
        img = res["image"].cpu().numpy()[:, 0]
        for n, (j, i) in enumerate(items):
            c = cs[n]
            s = nrm.offset_points_by_half_pixel(to_sample(c))
            want, _ = geo.focus_roi(s, geo.RoiFocusParams(c["scale"], c["angle"], c["translation"], host_cos_sin(c["angle"])),
                                    c["out_size"], use_model=True, upfilter=kind)
            check(img[n], gold["focus_" + kind][j], want.data["image"][0], kind == "cubic" and i == 31 and bool(gold["ipp"]), f"case {i}")


@pytest.mark.parametrize("kind", uc.FILTERS)
def test_tensor_entries_with_upfilter(gold, kind):
    from oracle import geometric as geo
    from trackertraincode_b200.datatransformation import tensors as dtt
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    for j, (i, entry, out_wh, g) in enumerate(uc.TENSOR_CASES):
        frame = cases.make_case(i)["image"]
        img = torch.from_numpy(frame[None].copy()).cuda()
        if entry == "crop":
            out = dtt.croprescale_image_cv2(img, torch.tensor(g, dtype=torch.int32), out_wh, upfilter=kind)
            model = geo.croprescale_image(frame, g, out_wh, use_model=True, upfilter=kind)
        else:
            tr = gold["tensor_tr"][j]
            out = dtt.affine_transform_image_cv2(img, Affine2d(torch.from_numpy(tr.copy())), out_wh, upfilter=kind)
            model = geo.affine_transform_image(frame, tr, out_wh, use_model=True, upfilter=kind)
        check(out.cpu().numpy()[0], gold[f"tensor_{kind}_{j}"], model, kind == "cubic" and entry == "crop" and bool(gold["ipp"]), f"tensor case {j}")


def test_random_upscaling_against_the_models():
    """Random boxes / transforms that grow, both up-filters together with a down-filter in one call, kernel vs oracle."""
    from oracle import geometric as geo
    from trackertraincode_b200.datatransformation import tensors as dtt
    from trackertraincode_b200.neuralnets.affine2d import Affine2d

    rng = np.random.default_rng(9)
    for t in range(20):
        h, w = (int(v) for v in rng.integers(40, 200, 2))
        frame = rng.integers(0, 256, (h, w), dtype=np.uint8)
        img = torch.from_numpy(frame[None].copy()).cuda()
        ow, oh = (int(v) for v in rng.integers(100, 260, 2))
        kind, down = uc.FILTERS[t % 2], ("area", "gaussian", "hamming")[t % 3]
        if t % 3:
            x0, y0 = (int(v) for v in rng.integers(-20, w // 2, 2))
            x1, y1 = x0 + int(rng.integers(8, 90)), y0 + int(rng.integers(8, 90))
            out = dtt.croprescale_image_cv2(img, torch.tensor((x0, y0, x1, y1), dtype=torch.int32), (ow, oh), downfilter=down, upfilter=kind)
            model = geo.croprescale_image(frame, (x0, y0, x1, y1), (ow, oh), use_model=True, downfilter=down, upfilter=kind)
        else:
            base = Affine2d.range_remap_2d([0.0, 0.0], [float(w), float(h)], [0.0, 0.0], [float(ow), float(oh)])
            tr = Affine2d.trs(translations=torch.tensor([float(rng.uniform(-9, 9)), float(rng.uniform(-9, 9))]),
                              angles=torch.tensor(float(rng.uniform(-0.6, 0.6))), scales=torch.tensor(float(rng.uniform(1.0, 2.0)))) @ base
            out = dtt.affine_transform_image_cv2(img, tr, (ow, oh), downfilter=down, upfilter=kind)
            model = geo.affine_transform_image(frame, tr.tensor().numpy(), (ow, oh), use_model=True, downfilter=down, upfilter=kind)
        assert np.array_equal(out.cpu().numpy()[0], model), (t, kind, down)


def test_full_chain_with_filters_on_a_mixed_batch():
    """One fused call -- half-pixel, focus with downfilter='gaussian' / 'hamming' and upfilter='lanczos', flip / rot90, normalise,
    photometric chain, whiten -- on a batch that mixes shrinking, growing, rotated and border-crossing crops, against the
    oracle's full chain sample by sample."""
    import bench
    from oracle import photometric as opho, pipeline as opipe
    from oracle.geometric import Sample
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import _engine as E

    B, S = 48, 129
    host = bench.make_host_batch(21, B)
    rng = np.random.default_rng(4)
    small = np.arange(B) % 3 == 0  # tiny faces: the crop grows (up-filter); rotated ones take the up-scaling warp
    c = 0.5 * (host["roi"][:, :2] + host["roi"][:, 2:])
    host["roi"][small] = np.concatenate([c[small] - rng.uniform(25, 45, (int(small.sum()), 2)), c[small] + rng.uniform(25, 45, (int(small.sum()), 2))], 1).astype(np.float32)
    host["roi"][1::7, [0, 2]] -= 150.0  # boxes over the left frame border
    gp, pp = bench.draw_params(77, B, 9000)
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    dev = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
    flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN
    an = torch.from_numpy(gp.angles)
    photo = E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma), torch.from_numpy(pp.contrast),
                          torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply), pp.noise_std, pp.seed, pp.sample_offset, pp.clip)
    for down in ("gaussian", "hamming"):
        b = Batch(Metadata((bench.SRC, bench.SRC), B, "t", None, dict(cats)), dict(dev))
        geo = E.GeoParams(torch.from_numpy(gp.scales), an, torch.from_numpy(gp.translations), E.host_cos_sin(an))
        r = E.fused_forward(b, flags=flags, out_size=S, geo=geo, do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)),
                            rot_dir=torch.from_numpy(gp.rot_dir), photo=photo, want_status=True, downfilter=down, upfilter="lanczos")
        assert not r.status.cpu().numpy().any()
        img = r.batch["image"].cpu().numpy()
        pts = r.batch["pt3d_68"].cpu().numpy()
        grew = 0
        for i in range(B):
            s1 = [Sample((bench.SRC, bench.SRC), {k: (host[k][i][..., None] if k == "image" else host[k][i]) for k in bench.CATS}, bench.CATS)]
            g1 = opipe.GeoParams(*(x[i:i + 1] for x in (gp.scales, gp.angles, gp.translations, gp.do_flip, gp.rot_dir)))
            want, inter = opipe.augment_batch(s1, g1, pp.slice(i, i + 1), S, use_model=True, downfilter=down, upfilter="lanczos")
            v = inter["view_roi"][0]
            grew += int(v[2] - v[0] < S)
            err = np.abs(img[i] - want["image"][0]).max()
            assert err <= 1e-5, (down, i, err)  # (kernel == oracle models, incl. the float32 hamming path)
            np.testing.assert_allclose(pts[i], want["pt3d_68"][0], rtol=1e-4, atol=2e-5)
        assert grew >= 8  # the up-filter really ran
