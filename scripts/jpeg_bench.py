#!/usr/bin/env python
"""Decode throughput of the JPEG -> grayscale frame entry beside the reference's decoder on the host:
512 colour JPEGs (450 x 450, quality 95, like the 300W-LP HDF5 blobs) per batch.   python scripts/jpeg_bench.py"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cv2  # noqa: E402
import numpy as np  # noqa: E402

import cases  # noqa: E402

B = 512


def _decode_chunk(blobs):
    cv2.setNumThreads(1)
    return sum(int(cv2.imdecode(b, 0)[0, 0]) for b in blobs)


def main():
    rng = np.random.default_rng(0)
    blobs = []
    for i in range(B):
        img = cases.make_image(rng, 450, 450, "smooth")
        img = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
        bgr = np.stack([img, np.roll(img, 5, 1), np.roll(img, 9, 0)], -1)
        blobs.append(cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, 95])[1].reshape(-1))
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    _decode_chunk(blobs[:128])
    one_core = 128 / (time.perf_counter() - t0)
    with ProcessPoolExecutor(cores) as ex:
        chunks = [blobs[i::cores] for i in range(cores)]
        list(ex.map(_decode_chunk, chunks))  # warm
        t0 = time.perf_counter()
        for _ in range(3):
            list(ex.map(_decode_chunk, chunks))
        all_cores = 3 * B / (time.perf_counter() - t0)

    import torch
    from trackertraincode_b200.datasets import preprocessing as pre

    for _ in range(3):
        pre.imdecode_batch(blobs, stack=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        pre.imdecode_batch(blobs, stack=True)
    torch.cuda.synchronize()
    gpu = n * B / (time.perf_counter() - t0)
    from trackertraincode_b200 import _native as N

    backend = {0: "none", 1: "nvJPEG GPU_HYBRID", 2: "nvJPEG HARDWARE (NVJPG engines)"}[int(N.lib.b200aug_jpeg_backend())]
    # pipelined: sub-batches on alternating streams, so that the host-side part of one call overlaps the device part of the other
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    sub = [blobs[i::4] for i in range(4)]
    t0 = time.perf_counter()
    for it in range(n):
        for k, sb in enumerate(sub):
            with torch.cuda.stream(streams[k % 2]):
                pre.imdecode_batch(sb, stack=True)
    torch.cuda.synchronize()
    gpu_piped = n * B / (time.perf_counter() - t0)
    print(json.dumps({"backend": backend, "gpu_images_per_s_4_subbatches_2_streams": gpu_piped, "workload": f"{B} colour JPEGs 450x450 q95 -> grayscale frames, mean blob {np.mean([len(b) for b in blobs]) / 1e3:.1f} KB",
                      "gpu_images_per_s": gpu, "cv2_one_core_images_per_s": one_core, "cv2_all_cores_images_per_s": all_cores, "cores": cores}))


if __name__ == "__main__":
    main()
