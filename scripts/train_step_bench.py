#!/usr/bin/env python
"""BASELINE.json configs[2]: one pose-net training step, batch 128, with the B200 augmentation feeding cuDNN fwd/bwd.

The CNN is NOT part of this repository's scope (it stays on PyTorch/cuDNN, north_star): a stock torchvision ResNet-18 with a
1-channel stem and the pose-net's output sizes (quaternion 4 + coord 3 + box 4 + 68x3 landmarks + 50 shape parameters,
neuralnets/models.py:244-330) stands in for `--backbone resnet18`, random init, synthetic data, plain L2 losses.  Measured:
  aug        FusedPoseAugmentation from pinned host frames (row-band upload + one fused launch)
  train      forward + backward + AdamW step on augmented crops already on the device
  step       both, as a training loop runs them: augmentation of batch i+1 on a side stream while batch i trains
and the CPU alternative (the oracle port of the reference chain on all host cores) for the same 128 samples.
Prints one JSON line; `python scripts/train_step_bench.py [--steps 50] [--amp]`.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

B = 128
N_OUT = 4 + 3 + 4 + 68 * 3 + 50


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the CNN")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    hosts = [bench.make_host_batch(7 + r, B) for r in range(2)]
    cpu = None
    if not args.no_cpu:
        cpu = bench.run_cpu_baseline(bench.make_host_batch(0), *bench.draw_params(100, bench.BATCH, 0), budget_s=4.0)

    import torchvision
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    net = torchvision.models.resnet18(num_classes=N_OUT)
    net.conv1 = torch.nn.Conv2d(1, 64, 7, 2, 3, bias=False)
    net = net.to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    pinned = [Batch(Metadata((bench.SRC, bench.SRC), B, "train", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in h.items()})
              for h in hosts]
    aug = FusedPoseAugmentation(bench.OUT, rotation_aug_angle=30.0, roi_override="original", enable_image_aug=True, device=dev)

    def train(batch):
        tgt = torch.cat([batch["pose"], batch["coord"], batch["roi"], batch["pt3d_68"].flatten(1),
                         torch.zeros(B, 50, device=dev)], 1)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp):
            out = net(batch["image"].contiguous(memory_format=torch.channels_last))
            loss = torch.nn.functional.mse_loss(out.float(), tgt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def timed(fn, n):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    ready = aug(pinned[0])
    ms_aug = timed(lambda i: aug(pinned[i % 2]), args.steps)
    ms_train = timed(lambda i: train(ready), args.steps)
    ms_serial = timed(lambda i: train(aug(pinned[i % 2])), args.steps)
    side = torch.cuda.Stream(dev)
    state = {"next": aug(pinned[0])}

    def overlapped(i):
        cur = state["next"]
        torch.cuda.current_stream().wait_stream(side)   # batch i is ready
        with torch.cuda.stream(side):                    # augment batch i+1 while batch i trains
            state["next"] = aug(pinned[(i + 1) % 2])
        train(cur)

    ms_step = timed(overlapped, args.steps)
    line = {"config": "configs[2]: pose-net training step, ResNet-18-class backbone, batch 128, B200 augmentation feeding cuDNN fwd/bwd, 1 B200",
            "batch": B, "amp_bf16": bool(args.amp), "steps": args.steps,
            "ms_aug_from_pinned_host": ms_aug, "ms_train_only": ms_train, "ms_step_serial": ms_serial, "ms_step_overlapped": ms_step,
            "aug_share_of_serial_step": ms_aug / ms_serial, "overhead_of_aug_when_overlapped_ms": ms_step - ms_train,
            "samples_per_s_step": B / ms_step * 1e3,
            "cpu_aug": None if cpu is None else {"samples_per_s": cpu["value"], "cores": cpu["cores"], "ms_per_128": B / cpu["value"] * 1e3,
                                                 "kind": cpu["kind"]}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
