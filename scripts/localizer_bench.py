#!/usr/bin/env python
"""BASELINE.json configs[4] shape (localizer): 640 x 480 gray frames, ROI crop + flip + photometric + bbox label, output
288 x 224 (LocalizerNet.input_resolution, neuralnets/models.py:33), batch 256, sources resident in HBM.  The reference has no
working localizer pipeline; this measures the same primitives (a7 / a12 / a15 / a17 / a20-22) through the fused
kernel.  Prints one JSON line.   python scripts/localizer_bench.py [--steps 100]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

B, W, H, OW, OH, RING = 256, 640, 480, 288, 224, 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    args = ap.parse_args()
    from oracle import geometric as ogeo
    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import _engine as E

    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:  # per-sample sharding, one process per GPU, no collective on the data path (weak scaling: 256 frames per GPU)
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    flags = N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN
    calls, alg = [], []
    for r in range(RING):
        rng = np.random.default_rng(50 + r)
        yy, xx = np.mgrid[0:H, 0:W]
        base = ((np.sin(xx / 17.0) + np.cos(yy / 23.0) + 2.0) / 4.0 * 255.0)
        img = np.clip(base[None] + rng.normal(0, 8, (B, H, W)), 0, 255).astype(np.uint8)
        wh = rng.uniform(150, 330, (B, 2))
        c = np.stack([rng.uniform(200, 440, B), rng.uniform(150, 330, B)], 1)
        roi = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        gp, pp = bench.draw_params(300 + r, B, r * B)
        gp.angles[:] = 0  # the localizer crop is not rotated
        gp.rot_dir[:] = 0
        cats = {"image": FieldCategory.image, "roi": FieldCategory.roi}
        batch = Batch(Metadata((W, H), B, "loc", None, cats), {"image": torch.from_numpy(img).to(dev), "roi": torch.from_numpy(roi).to(dev)})
        photo = E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma), torch.from_numpy(pp.contrast),
                              torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply), pp.noise_std, pp.seed, pp.sample_offset, pp.clip)
        an = torch.from_numpy(gp.angles)
        geo = E.GeoParams(torch.from_numpy(gp.scales), an, torch.from_numpy(gp.translations), E.host_cos_sin(an))
        calls.append(E.prepare_fused(batch, flags=flags, out_size=(OW, OH), geo=geo, do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)),
                                     rot_dir=None, photo=photo, want_status=True))  # (90-degree rotations need a square output)
        v = ogeo.round_view_roi(ogeo.compute_view_roi(roi, gp.scales, gp.translations)).astype(np.int64)
        w_ = np.clip(np.minimum(v[:, 2], W) - np.maximum(v[:, 0], 0), 0, None)
        h_ = np.clip(np.minimum(v[:, 3], H) - np.maximum(v[:, 1], 0), 0, None)
        alg.append(int((w_ * h_).sum() + B * (OW * OH * 4 + 2 * 16)))
    for i in range(20):
        calls[i % RING].launch()
    torch.cuda.synchronize()
    for c_ in calls:
        st = c_.result.status.cpu().numpy()
        assert not st.any(), f"per-sample status {np.unique(st)}"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for s in range(args.steps):
        calls[s % RING].launch()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.steps * 1e3
    if world > 1:
        t = torch.tensor([us], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
        dist.destroy_process_group()
        if rank != 0:
            return
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    gbs = float(np.mean(alg)) / (us * 1e-6) / 1e9
    print(json.dumps({"config": f"configs[4] shape: {B} x {W}x{H} u8 gray -> {OW}x{OH} f32, ROI crop + flip + photometric + whiten + roi label, ring of "
                                f"{RING} batches ({RING * B * W * H / 1e6:.0f} MB of sources), {world} x B200 (256 frames per GPU, no collective)",
                      "n_gpus": world, "kernel_us": us, "samples_per_s": world * B / (us * 1e-6), "algorithmic_bytes_per_launch": float(np.mean(alg)),
                      "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak}))


if __name__ == "__main__":
    main()
