#!/usr/bin/env python
"""BASELINE.json configs[2] / configs[3]: the pose-net training step with the B200 augmentation feeding cuDNN forward /
backward -- one process per GPU, per-rank sharded augmentation (no collective on the data path), DDP gradient all-reduce
over NCCL (the only collective, exactly as in the reference: scripts/train_poseestimator.py:310-330 training_step,
:442-454 the Trainer that wraps the model in DDP).

The CNN is NOT this repository's product (north_star: it stays on PyTorch / cuDNN): a stock torchvision ResNet-18 with a
1-channel stem and the pose-net's output sizes (quaternion 4 + coord 3 + box 4 + 68 x 3 landmarks + 50 shape parameters,
neuralnets/models.py:244-330) stands in for `--backbone resnet18`; random init, synthetic frames, plain L2 losses, AdamW,
gradient-norm clipping at 1.0 (train_poseestimator.py:444-445).

Per step and rank: FusedPoseAugmentation of batch i + 1 from pinned host frames on a side stream (host parameter sampling,
row-band upload, plan + fused kernels) while batch i trains (forward, backward with bucketed all-reduce, clip, optimizer).

  python scripts/train_bench.py [--batch 128] [--steps 40]                      config 3 (1 GPU, batch 128)
  torchrun --nproc-per-node N scripts/train_bench.py --batch 256 [--steps 40]   config 4 (N x 256)
or through bench.py: `python bench.py --workload train --gpus N`.  Prints one JSON line (rank 0)."""
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

N_OUT = 4 + 3 + 4 + 68 * 3 + 50


def run(batch: int, steps: int, warmup: int, amp: bool, cpu_ref: bool = True):
    import bench

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    hosts = [bench.make_host_batch(7 + r, batch) for r in range(2)]
    cpu = None
    if cpu_ref and rank == 0 and world == 1:
        cpu = bench.run_cpu_baseline(bench.make_host_batch(0), *bench.draw_params(100, bench.BATCH, 0), budget_s=4.0)

    import torch
    import torch.distributed as dist
    import torchvision
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation

    if world > 1:
        bench.pin_to_gpu_cores(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)
    net = torchvision.models.resnet18(num_classes=N_OUT)
    net.conv1 = torch.nn.Conv2d(1, 64, 7, 2, 3, bias=False)
    net = net.to(dev).to(memory_format=torch.channels_last).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local]) if world > 1 else net
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
    cats = {k: FieldCategory(v) for k, v in bench.CATS.items()}
    pinned = [Batch(Metadata((bench.SRC, bench.SRC), batch, "train", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in h.items()})
              for h in hosts]
    aug = FusedPoseAugmentation(bench.OUT, rotation_aug_angle=30.0, roi_override="original", enable_image_aug=True, device=dev)
    zeros50 = torch.zeros(batch, 50, device=dev)

    def train(b):
        tgt = torch.cat([b["pose"], b["coord"], b["roi"], b["pt3d_68"].flatten(1), zeros50], 1)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            out = model(b["image"].contiguous(memory_format=torch.channels_last))
            loss = torch.nn.functional.mse_loss(out.float(), tgt)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        """ms per step: CUDA events on the current stream around n steps, max over ranks."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) / n * 1e3
        ms = max(e0.elapsed_time(e1) / n, 0.0)
        if world > 1:
            t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = (float(x) for x in t.cpu())
        return ms, wall

    ready = aug(pinned[0])
    ms_aug, _ = timed(lambda i: aug(pinned[i % 2]), steps)
    ms_train, _ = timed(lambda i: train(ready), steps)
    side = torch.cuda.Stream(dev)
    state = {"next": aug(pinned[0])}

    def overlapped(i):
        cur = state["next"]
        torch.cuda.current_stream().wait_stream(side)   # batch i is augmented
        with torch.cuda.stream(side):                    # augment batch i + 1 while batch i trains
            side.wait_stream(torch.cuda.current_stream())  # (its output buffers are new tensors; this only orders the launch)
            state["next"] = aug(pinned[(i + 1) % 2])
        loss = train(cur)
        if i % 10 == 9:
            loss.item()  # a training loop reads the loss now and then

    ms_step, wall_step = timed(overlapped, steps)
    aug.status.flush()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return None
    return {
        "workload": ("configs[3]: DDP training, per-rank sharded GPU augmentation" if world > 1 else "configs[2]: pose-net training step") +
                    f", ResNet-18-class backbone, batch {batch} per GPU, B200 augmentation feeding cuDNN fwd/bwd, {world} x B200",
        "n_gpus": world, "batch_per_gpu": batch, "global_batch": world * batch, "amp_bf16": bool(amp), "steps": steps,
        "samples_per_s": world * batch / max(ms_step, wall_step) * 1e3, "ms_per_step": max(ms_step, wall_step),
        "ms_step_device": ms_step, "ms_step_wall": wall_step, "ms_train_only": ms_train, "ms_aug_from_pinned_host": ms_aug,
        "aug_share_if_serial": ms_aug / (ms_aug + ms_train), "overhead_of_aug_when_overlapped_ms": max(ms_step, wall_step) - ms_train,
        "collective": "DDP gradient all-reduce only (NCCL); the augmentation has none",
        "cpu_aug": None if cpu is None else {"samples_per_s": cpu["value"], "cores": cpu["cores"], "ms_per_batch": batch / cpu["value"] * 1e3,
                                             "kind": cpu["kind"]},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch; 0 = 128 on one GPU (config 3), 256 under torchrun (config 4)")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast for the CNN")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = run(args.batch or (256 if world > 1 else 128), args.steps, args.warmup, args.amp, not args.no_cpu)
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
