#!/usr/bin/env python
"""Where the fused kernel's DRAM reads go: the config-2 batch launched three times under ncu --
(1) as drawn, (2) the same draws with every angle set to 0 (no rotated sample), (3) the photometric stage off -- next to the
source bytes each variant needs by construction:
  view boxes ∩ frame (unrotated samples), bounding boxes of the rotated squares ∩ frame (rotated samples: cv2.warpAffine
  reads the rotated square, the staging fetches its bounding box), and both again in 32-byte sectors per row.

Run:  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,\
lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:fused_augment --csv --log-file out.csv \
python scripts/traffic_attribution.py     (the script prints the by-construction table as JSON; launches appear in order)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench  # noqa: E402
from oracle import geometric as ogeo  # noqa: E402
from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata  # noqa: E402
from trackertraincode_b200.datatransformation import FusedPoseAugmentation, _engine as E  # noqa: E402

B, S = bench.BATCH, bench.SRC


def source_bytes(roi, scales, translations, angles):
    v = ogeo.round_view_roi(ogeo.compute_view_roi(roi, scales, translations)).astype(np.float64)
    cx, cy = 0.5 * (v[:, 0] + v[:, 2]), 0.5 * (v[:, 1] + v[:, 3])
    side = v[:, 2] - v[:, 0]
    a = np.abs(angles.astype(np.float64))
    half = 0.5 * side * (np.cos(a) + np.sin(a))  # (= side / 2 for angle 0)
    x0, x1 = np.clip(np.floor(cx - half), 0, S), np.clip(np.ceil(cx + half), 0, S)
    y0, y1 = np.clip(np.floor(cy - half), 0, S), np.clip(np.ceil(cy + half), 0, S)
    w, h = np.clip(x1 - x0, 0, None), np.clip(y1 - y0, 0, None)
    # 32-byte sectors a row segment [x0, x1) of a frame with pitch S touches, averaged over the row's alignment
    sect = np.zeros_like(w)
    for i in range(len(w)):
        if w[i] > 0 and h[i] > 0:
            rows = np.arange(int(y0[i]), int(y1[i]))
            start = rows * S + int(x0[i])  # (frame base is 256-byte aligned; frame stride S * S is not a multiple of 32 -- ignored)
            sect[i] = ((start + int(w[i]) - 1) // 32 - start // 32 + 1).sum() * 32
    rot = angles != 0
    return dict(unrotated_box_bytes=int((w * h)[~rot].sum()), rotated_bbox_bytes=int((w * h)[rot].sum()),
                unrotated_sector_bytes=int(sect[~rot].sum()), rotated_sector_bytes=int(sect[rot].sum()),
                rotated_samples=int(rot.sum()), view_box_bytes_SURVEY_8d=int(np.clip(np.minimum(v[:, 2], S) - np.maximum(v[:, 0], 0), 0, None)
                                                                            @ np.clip(np.minimum(v[:, 3], S) - np.maximum(v[:, 1], 0), 0, None)))


def main():
    dev = torch.device("cuda")
    host = bench.make_host_batch(0)
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    batch = Batch(Metadata((S, S), B, "t", None, dict(cats)), {k: torch.from_numpy(v).to(dev) for k, v in host.items()})
    aug = FusedPoseAugmentation(bench.OUT, rotation_aug_angle=30.0, device=dev)
    torch.manual_seed(0)
    np.random.seed(0)
    d = aug.draw(B)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    table = {}
    for name in ("as_drawn", "no_rotation", "no_photometric"):
        if name == "no_rotation":
            d.geo = E.GeoParams(d.geo.scales, torch.zeros_like(d.geo.angles), d.geo.translations, E.host_cos_sin(torch.zeros_like(d.geo.angles)))
        if name == "no_photometric":
            aug.flags &= ~0x20
            d.photo = None
        flush.fill_(1)  # cold L2, like the first launch of a step on fresh frames
        torch.cuda.synchronize()
        aug(batch, params=d)
        torch.cuda.synchronize()
        table[name] = source_bytes(host["roi"], d.geo.scales.numpy(), d.geo.translations.numpy(), d.geo.angles.numpy())
        table[name]["output_bytes"] = B * bench.OUT * bench.OUT * 4
        table[name]["plan_record_bytes"] = int(B * E.N.lib.b200aug_plan_stride(bench.OUT, bench.OUT))
    print(json.dumps(table))


if __name__ == "__main__":
    main()
