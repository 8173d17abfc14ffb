#!/usr/bin/env python
"""Host-side cost of one FusedPoseAugmentation call (cProfile over 100 calls, batch 512 from pinned host frames)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata  # noqa: E402
from trackertraincode_b200.datatransformation import FusedPoseAugmentation  # noqa: E402

B = bench.BATCH
h = bench.make_host_batch(0)
cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
pinned = Batch(Metadata((bench.SRC, bench.SRC), B, "p", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in h.items()})
aug = FusedPoseAugmentation(bench.OUT, device="cuda")
for _ in range(10):
    aug(pinned)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    aug(pinned)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
