#!/usr/bin/env python
"""Host-side cost of one FusedPoseAugmentation call in the steady state of bench.py's e2e loop (two streams alternate, the
host waits for step s - 1 after enqueuing step s): cProfile over 100 steps, batch 512 from pinned host frames."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata  # noqa: E402
from trackertraincode_b200.datatransformation import FusedPoseAugmentation  # noqa: E402

B = bench.BATCH
cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
pinned = [Batch(Metadata((bench.SRC, bench.SRC), B, "p", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in bench.make_host_batch(i).items()})
          for i in range(2)]
aug = FusedPoseAugmentation(bench.OUT, device="cuda")
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
keys = ("roi", "coord", "pose", "pt3d_68")
outs = [{k: torch.empty_like(pinned[0][k]).pin_memory() for k in keys} for _ in range(2)]


def step(s):
    with torch.cuda.stream(streams[s % 2]):
        out = aug(pinned[s % 2])
        for k in keys:
            outs[s % 2][k].copy_(out[k], non_blocking=True)
    streams[(s - 1) % 2].synchronize()


for s in range(10):
    step(s)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in range(100):
    step(s)
pr.disable()
torch.cuda.synchronize()
print(f"{(time.perf_counter() - t0) * 10:.3f} ms per step")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
