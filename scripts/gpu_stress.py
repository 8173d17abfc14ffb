#!/usr/bin/env python
"""Randomised parity sweep (GPU box): FusedPoseAugmentation on ragged batches of odd frame sizes / pitches, rotation angles up
to 45 degrees, several output sizes, against the oracle with identical draws.  Prints the worst errors; exit code 1 on a
violation of the tolerances (pixels 1/255, labels 1e-4 relative).   python scripts/gpu_stress.py [rounds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402
from oracle import photometric as opho, pipeline as opipe  # noqa: E402
from oracle.geometric import Sample  # noqa: E402
from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata  # noqa: E402
from trackertraincode_b200.datatransformation import FusedPoseAugmentation  # noqa: E402

CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")
SIZES = [(450, 450), (451, 333), (640, 480), (200, 180), (97, 131), (333, 517)]


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    worst_px, worst_lab, bad = 0.0, 0.0, 0
    for rnd in range(rounds):
        rng = np.random.default_rng(1000 + rnd)
        S = [129, 64, 200, 129][rnd % 4]
        n = 48
        ragged = rnd % 3 != 0
        sizes = [SIZES[int(rng.integers(len(SIZES)))] if ragged else SIZES[rnd % len(SIZES)] for _ in range(n)]
        pairs = []
        for i in range(n):
            w, h = sizes[i]
            lab = cases.make_labels(rng, w, h)
            lab.pop("shapeparam")
            raw = dict(image=cases.make_image(rng, w, h, "noise" if i % 2 else "smooth"), **lab)
            meta = Metadata((w, h), 0, "s", None, {k: FieldCategory(v) for k, v in CATS.items()})
            pairs.append((Batch(meta, {k: torch.from_numpy(v) for k, v in raw.items()}), raw))
        collated = Batch.Collation(ragged_images=ragged)([p[0] for p in pairs])
        aug = FusedPoseAugmentation(S, rotation_aug_angle=[30.0, 45.0, 10.0][rnd % 3], device="cuda", seed=rnd)
        torch.manual_seed(rnd)
        np.random.seed(rnd)
        d = aug.draw(n)
        out = aug(collated.pin_memory(), params=d)
        gp = opipe.GeoParams(d.geo.scales.numpy(), d.geo.angles.numpy(), d.geo.translations.numpy(), d.do_flip.numpy().astype(bool), d.rot_dir.numpy())
        p = d.photo
        pp = opho.PhotoParams(list(p.order), p.apply.numpy(), p.bits.numpy(), p.gamma.numpy(), p.contrast.numpy(), p.brightness.numpy(),
                              p.noise_apply.numpy(), p.noise_std, p.seed, p.sample_offset, p.clip)
        samples = [Sample(sizes[i], {k: (v[..., None] if k == "image" else v) for k, v in raw.items()}, CATS) for i, (_, raw) in enumerate(pairs)]
        want, _ = opipe.augment_batch(samples, gp, pp, S)
        e_px = float(np.abs(out["image"].cpu().numpy() - want["image"]).max())
        e_lab = 0.0
        for k in ("roi", "coord", "pt3d_68"):
            g, w_ = out[k].cpu().numpy(), want[k]
            e_lab = max(e_lab, float((np.abs(g - w_) / (1e-4 * np.abs(w_) + 2e-5)).max()))
        q, qw = out["pose"].cpu().numpy(), want["pose"]
        e_lab = max(e_lab, float(np.minimum(np.abs(q - qw).max(-1), np.abs(q + qw).max(-1)).max() / 2e-5))
        ok = e_px <= 1.0 / 255 and e_lab <= 1.0
        bad += not ok
        worst_px, worst_lab = max(worst_px, e_px), max(worst_lab, e_lab)
        print(f"round {rnd}: S={S} ragged={ragged} rot<={[30, 45, 10][rnd % 3]} rotated={int((gp.angles != 0).sum())} px_err={e_px:.3g} "
              f"label_err/tol={e_lab:.3g} {'ok' if ok else 'VIOLATION'}", flush=True)
    print(f"worst pixel error {worst_px:.3g} (tol {1 / 255:.3g}); worst label error / tolerance {worst_lab:.3g}; violations {bad}/{rounds}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
