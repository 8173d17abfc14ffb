#!/usr/bin/env python
"""The loader-side drop-in with REAL worker processes (pipelines.py:359-554): a dataset whose workers
return the raw sample -- (a) the decoded 450 x 450 uint8 frame, or (b) the still encoded JPEG blob, as the HDF5
`varsize_image_buffer` stores it (datasets/dshdf5.py:59-113) -- through `SegmentedCollationDataLoader(pin_memory=True)` into `FusedPoseAugmentation` installed as `postprocess` (blobs are decoded on the GPU first:
`datasets.preprocessing.imdecode_batch`, nvJPEG).  Measures frames/s of the whole loader loop and the PCIe bytes per frame.
For comparison the reference arrangement: workers run the CPU chain (oracle port) and return 129 x 129 crops.

  python scripts/loader_bench.py [--workers 12] [--batches 40] [--batch 256]     prints one JSON line"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cases  # noqa: E402

CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")
N_FRAMES = 64  # distinct frames the synthetic dataset cycles through (the bundled aflw2kmini.h5 has 16)


class RawDataset(torch.utils.data.Dataset):
    """What Hdf5PoseDataset.__getitem__ returns when the per-sample transform chain is empty: the frame (decoded, or as the
    JPEG blob) + labels, one `Batch` per index."""

    def __init__(self, n: int, mode: str):
        self.n, self.mode = n, mode
        rng = np.random.default_rng(0)
        self.frames, self.blobs, self.labels = [], [], []
        for i in range(N_FRAMES):
            img = cases.make_image(rng, 450, 450, "smooth")
            bgr = np.stack([img, np.roll(img, 5, 1), np.roll(img, 9, 0)], -1)
            blob = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].reshape(-1)
            lab = cases.make_labels(rng, 450, 450)
            lab.pop("shapeparam")
            self.frames.append(img)
            self.blobs.append(blob)
            self.labels.append(lab)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata

        j = i % N_FRAMES
        lab = {k: torch.from_numpy(v) for k, v in self.labels[j].items()}
        if self.mode == "jpeg":
            image = torch.from_numpy(self.blobs[j])                      # what the HDF5 file holds: no decode in the worker
        elif self.mode == "decode":
            image = torch.from_numpy(cv2.imdecode(self.blobs[j], 0))     # the reference's worker: cv2.imdecode(blob, 0)
        else:
            image = torch.from_numpy(self.frames[j])                     # frames already decoded (cached dataset)
        return Batch(Metadata((450, 450), 0, "ds", None, {k: FieldCategory(v) for k, v in CATS.items()}), {"image": image, **lab})


def run(mode: str, workers: int, batches: int, batch: int):
    from trackertraincode_b200.datasets import preprocessing as pre
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation, SegmentedCollationDataLoader

    dev = torch.device("cuda", 0)
    aug = FusedPoseAugmentation(129, rotation_aug_angle=30.0, device=dev)
    h2d = [0]

    def postprocess(b):
        if mode == "jpeg":
            blobs = b["image"]
            h2d[0] += sum(int(x.numel()) for x in blobs)
            b = b.__class__(b.meta, {**dict(b.items()), "image": pre.imdecode_batch([x.numpy() for x in blobs], device=dev, stack=True)})
            return aug(b.to(dev, non_blocking=True))
        out = aug(b)  # stacked [B, 450, 450] uint8, pinned by the loader's pin_memory thread: row-band upload + one launch
        h2d[0] += aug.uploaded_bytes
        return out

    ds = RawDataset(batch * (batches + 4), mode)
    # equal-sized frames are stacked by the worker's collation (one shared-memory tensor per batch); JPEG blobs have
    # different lengths and travel as a ragged list
    loader = SegmentedCollationDataLoader(ds, batch_size=batch, num_workers=workers, segmentation_key_getter=lambda s: s.meta.tag,
                                          pin_memory=True, postprocess=postprocess, ragged_images=(mode == "jpeg"))
    n, t0, out = 0, None, None
    for i, items in enumerate(loader):
        if i == 4:  # workers spun up, kernels warm
            torch.cuda.synchronize()
            t0, n, h2d[0] = time.perf_counter(), 0, 0
        for out in items:
            n += out.meta.batchsize
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert out["image"].shape[1:] == (1, 129, 129) and out["image"].dtype == torch.float32
    aug.status.flush()
    return {"frames_per_s": n / dt, "h2d_bytes_per_frame": h2d[0] / max(n, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=max(1, (os.cpu_count() or 2) - 4))
    ap.add_argument("--batches", type=int, default=40)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    res = {m: run(m, args.workers, args.batches, args.batch) for m in ("frames", "decode", "jpeg")}
    print(json.dumps({"loader": "SegmentedCollationDataLoader(pin_memory=True) -> postprocess = FusedPoseAugmentation",
                      "workers": args.workers, "batch": args.batch, "batches": args.batches,
                      "raw_frames_from_workers": res["frames"], "workers_decode_jpeg_cv2": res["decode"],
                      "jpeg_blobs_decoded_on_gpu_nvjpeg": res["jpeg"]}))


if __name__ == "__main__":
    main()
