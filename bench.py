#!/usr/bin/env python
"""Benchmark of the augmentation / label-transform hot path (BASELINE.json metric: augmented samples/s).

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  CPU implementation on the host cores

Workload (BASELINE.json configs[1], SURVEY.md 8d "config 2"): synthetic 300W-LP-shaped batch -- 450x450 uint8 gray
sources -> 129x129 float32 crops, batch 512 per GPU, full geometric (crop/scale/translate, 1/3 rotated by +-30 deg, flip,
rot90) + photometric (equalize/posterize/gamma/contrast/brightness/blur, 4 noise stages, clip) + whiten + all labels.
One step = one pass of the hot path over one batch.  Prints ONE JSON line (rank 0).

value      whole-job samples/s with sources and sampled parameters resident in HBM (the fused kernel alone)
e2e        samples/s through the public API (`FusedPoseAugmentation(batch)`) from pinned HOST buffers: per step the
           H2D copy of the uint8 frames + labels, host-side parameter sampling + upload, the kernel, and a D2H read of
           the transformed labels
roofline   algorithmic bytes (clipped view-box pixels x 1 B + output x 4 B + label bytes) / event-timed kernel duration
           against MEASURED_PEAKS.json hbm_gbs
cpu_baseline  the oracle port (numpy + cv2, one process per host core) on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

BATCH = 512
SRC = 450
OUT = 129
RING = 4  # distinct batches resident in HBM: 4 x 104 MB of sources > 126 MB L2
WORKLOAD = ("config2: synthetic 300W-LP-shaped, 450x450 u8 gray -> 129x129 f32, batch 512 per GPU, "
            "full geometric+photometric+labels")
LABEL_BYTES = 2 * (68 * 3 * 4 + 16 + 16 + 12) + 40


# ----------------------------------------------------------------------------------------------- synthetic workload

def make_host_batch(seed: int, n: int = BATCH):
    """Sources + labels of SURVEY.md 8d config 2, numpy on the host."""
    rng = np.random.default_rng(seed)
    img = np.empty((n, SRC, SRC), np.uint8)
    half = n // 2
    img[:half] = rng.integers(0, 256, (half, SRC, SRC), dtype=np.uint8)  # (i) uniform noise
    y, x = np.mgrid[0:SRC, 0:SRC].astype(np.float32)
    base = (np.sin(x / 17.0) + np.cos(y / 23.0) + 2.0) / 4.0 * 255.0  # (ii) smooth field + N(0, 8)
    for i in range(half, n):
        img[i] = np.clip(np.rint(base + rng.standard_normal((SRC, SRC), dtype=np.float32) * 8.0), 0, 255).astype(np.uint8)
    bw, bh = rng.uniform(147, 250, n), rng.uniform(147, 250, n)
    cx, cy = 225 + rng.uniform(-40, 40, n), 225 + rng.uniform(-40, 40, n)
    roi = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], -1).astype(np.float32)
    coord = np.stack([cx, cy, 0.5 * np.maximum(bw, bh)], -1).astype(np.float32)
    q = rng.standard_normal((n, 4))
    pose = (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32)
    pts = np.empty((n, 68, 3), np.float32)
    pts[..., 0] = rng.uniform(roi[:, None, 0], roi[:, None, 2], (n, 68))
    pts[..., 1] = rng.uniform(roi[:, None, 1], roi[:, None, 3], (n, 68))
    pts[..., 2] = rng.normal(0, 30, (n, 68))
    return dict(image=img, roi=roi, coord=coord, pose=pose, pt3d_68=pts)


def draw_params(seed: int, n: int, sample_offset: int):
    from oracle import photometric as opho, pipeline as opipe

    rng = np.random.default_rng(seed)
    gp, pp = opipe.sample_geo_params(rng, n), opho.sample_photo_params(rng, n, seed=5, sample_offset=sample_offset)
    # experiment switches for kernel work (never set for a reported number): which parts of the workload to drop
    drop = os.environ.get("B200AUG_BENCH_DROP", "").split(",")
    if "rot" in drop:
        gp.angles[:] = 0
    if "allrot" in drop:
        gp.angles[:] = np.float32(np.pi / 6)
    if "photo" in drop:
        pp.apply[:] = False
        pp.noise_apply[:] = False
    if "blur" in drop:
        pp.apply[:, 5] = False
    if "eq" in drop:
        pp.apply[:, 0] = False
    if "noise" in drop:
        pp.noise_apply[:] = False
    return gp, pp


def algorithmic_bytes(host, gp) -> int:
    """view-box pixels inside the frame (1 B each) + float32 output + labels, summed over the batch."""
    from oracle import geometric as ogeo

    v = ogeo.round_view_roi(ogeo.compute_view_roi(host["roi"], gp.scales, gp.translations)).astype(np.int64)
    w = np.clip(np.minimum(v[:, 2], SRC) - np.maximum(v[:, 0], 0), 0, None)
    h = np.clip(np.minimum(v[:, 3], SRC) - np.maximum(v[:, 1], 0), 0, None)
    return int((w * h).sum() + len(v) * (OUT * OUT * 4 + LABEL_BYTES))


# ----------------------------------------------------------------------------------------------- CPU baseline (oracle port)

_CPU_WORK = None
CATS = dict(image="img", roi="roi", coord="xys", pose="q", pt3d_68="pts")


def _cpu_init():
    import cv2

    cv2.setNumThreads(1)  # like the reference's DataLoader workers (pipelines.py:59-69)
    try:
        import torch

        torch.set_num_threads(1)
    except Exception:
        pass


def _cpu_chunk(rng_):
    lo, hi = rng_
    from oracle import pipeline as opipe
    from oracle.geometric import Sample

    host, gp, pp = _CPU_WORK
    n = len(host["roi"])
    idx = [i % n for i in range(lo, hi)]
    samples = [Sample((SRC, SRC), {k: (host[k][i][..., None] if k == "image" else host[k][i]) for k in CATS}, CATS) for i in idx]
    sel = lambda a: a[idx]  # noqa: E731
    from oracle.photometric import PhotoParams
    from oracle.pipeline import GeoParams

    g = GeoParams(sel(gp.scales), sel(gp.angles), sel(gp.translations), sel(gp.do_flip), sel(gp.rot_dir))
    p = PhotoParams(pp.order, sel(pp.apply), sel(pp.bits), sel(pp.gamma), sel(pp.contrast), sel(pp.brightness),
                    sel(pp.noise_apply), pp.noise_std, pp.seed, pp.sample_offset + lo, pp.clip)
    out, _ = opipe.augment_batch(samples, g, p, OUT)
    return float(out["image"].sum())


class CpuPool:
    """The oracle's full chain (per-sample geometric/label half + photometric half + whiten) on every host core."""

    CHUNK = 8

    def __init__(self, host, gp, pp):
        import multiprocessing as mp

        global _CPU_WORK
        _CPU_WORK = (host, gp, pp)
        self.cores = len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)
        self.pool.map(_cpu_chunk, [(i, i + 1) for i in range(self.cores)])  # spin the workers up

    def run(self, n_samples: int) -> float:
        """Process n_samples (cycling through the batch); returns wall seconds."""
        tasks = [(lo, min(lo + self.CHUNK, n_samples)) for lo in range(0, n_samples, self.CHUNK)]
        t0 = time.perf_counter()
        self.pool.map(_cpu_chunk, tasks, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_baseline(host, gp, pp, budget_s=12.0):
    pool = CpuPool(host, gp, pp)
    t = pool.run(pool.cores * CpuPool.CHUNK)  # calibration
    rate = pool.cores * CpuPool.CHUNK / t
    n = int(min(max(rate * budget_s, BATCH), 16 * BATCH) // CpuPool.CHUNK * CpuPool.CHUNK)
    t = pool.run(n)
    pool.close()
    return dict(value=n / t, unit="samples/s", cores=pool.cores, kind="port",
                sample=f"{n} samples of the workload (batch 0 cycled), oracle/pipeline.py full chain, {pool.cores} single-threaded "
                       f"processes, {t:.1f} s wall; CPU: {cpu_model_name()}")


def bench_config(world: int):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": world * BATCH}


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path (oracle port; the Python reference itself cannot travel to the
    GPU box) with every host core, one step = one 512-sample batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    host = make_host_batch(0)
    gp, pp = draw_params(100, BATCH, 0)
    pool = CpuPool(host, gp, pp)
    for _ in range(max(args.warmup, 1)):
        pool.run(BATCH)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(BATCH)
    dt = time.perf_counter() - t0
    pool.close()
    v = BATCH * args.steps / dt
    cb = dict(value=v, unit="samples/s", cores=pool.cores, kind="port",
              sample=f"{args.steps} steps x {BATCH} samples, oracle/pipeline.py full chain, {pool.cores} single-threaded processes; "
                     f"CPU: {cpu_model_name()}")
    print(json.dumps({
        "impl": "reference", "metric": "augmented_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "details": {"note": "runs on the host cores of rank 0 only; one step = one 512-sample batch"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------- GPU arm

class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, power) if p >= 0.6 * max(power)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(power))}


def pin_to_gpu_cores(gpu_index: int):
    """Bind this rank (and the threads it spawns) to the host cores NVML reports as local to its GPU (same NUMA node / PCIe
    root): eight ranks sampling parameters and driving 63 MB/step copies each otherwise wander over both sockets."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cores &= os.sched_getaffinity(0)
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:  # no NVML / not permitted: stay unpinned
        pass
    return None


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from trackertraincode_b200 import _native as N
    from trackertraincode_b200.datasets.batch import Batch, FieldCategory, Metadata
    from trackertraincode_b200.datatransformation import FusedPoseAugmentation, _engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    # host data + the CPU baseline first (forking a pool after CUDA is initialised is asking for trouble).
    # Weak scaling: every rank works on the SAME four batches and draws (per-GPU work is fixed as N grows; with per-rank
    # seeds the max over ranks measured which rank had drawn the dearest samples -- SCALE_r01: 0.86 "efficiency" without
    # any collective on the path); only the ids of the noise stream differ by rank.
    hosts = [make_host_batch(r) for r in range(RING)]
    params = [draw_params(100 + r, BATCH, (rank * RING + r) * BATCH) for r in range(RING)]
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(hosts[0], *params[0])
    affinity = pin_to_gpu_cores(local) if world > 1 and not args.no_pin else None

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cats = {k: FieldCategory(v) for k, v in CATS.items()}
    flags = N.F_HALF_PIXEL | N.F_FOCUS | N.F_FLIPROT | N.F_NORMALIZE | N.F_PHOTOMETRIC | N.F_WHITEN

    def device_batch(h):
        return Batch(Metadata((SRC, SRC), BATCH, "bench", None, dict(cats)), {k: torch.from_numpy(v).to(dev) for k, v in h.items()})

    def to_photo(pp):
        return E.PhotoParams(pp.order, torch.from_numpy(pp.apply), torch.from_numpy(pp.bits), torch.from_numpy(pp.gamma),
                             torch.from_numpy(pp.contrast), torch.from_numpy(pp.brightness), torch.from_numpy(pp.noise_apply),
                             pp.noise_std, pp.seed, pp.sample_offset, pp.clip)

    calls = []
    for h, (gp, pp) in zip(hosts, params):
        an = torch.from_numpy(gp.angles)
        geo = E.GeoParams(torch.from_numpy(gp.scales), an, torch.from_numpy(gp.translations), E.host_cos_sin(an))
        calls.append(E.prepare_fused(device_batch(h), flags=flags, out_size=OUT, geo=geo,
                                     do_flip=torch.from_numpy(gp.do_flip.astype(np.uint8)), rot_dir=torch.from_numpy(gp.rot_dir),
                                     photo=to_photo(pp), want_status=True, rowbuf_capacity=args.rowbuf, private_scratch=True,
                                     cluster_size=int(os.environ.get("B200AUG_BENCH_CLUSTER", "0"))))
    alg_bytes = float(np.mean([algorithmic_bytes(h, gp) for h, (gp, _) in zip(hosts, params)]))
    # The big kernel's stream gets a higher priority than the side stream of the plan phase: the block scheduler then hands a
    # freed slot to a pending CTA of the big kernel first, so plan_kernel(s + 1) only ever runs in the tail of step s (where
    # the slots are free anyway) instead of competing with the head of a step (measured: windows of 92.8 and of 102.6 us per
    # step in the same run, depending on where the plan grids happened to land).
    prio = os.environ.get("B200AUG_BENCH_PRIORITY", "1") != "0"
    stream = torch.cuda.Stream(dev, priority=-1) if prio else torch.cuda.current_stream(dev)
    if prio:
        stream.wait_stream(torch.cuda.current_stream(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- kernel-resident measurement ("value")
    sampler = ClockSampler(local) if rank == 0 else None
    t_spin = time.perf_counter()
    i = 0
    while i < args.warmup or time.perf_counter() - t_spin < args.spin_s:  # warm-up: >= W steps and long enough for clocks to settle
        calls[i % RING].launch()
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize(dev)
    for c in calls:
        assert not c.result.status.cpu().numpy().any(), "per-sample status reported a problem"
    barrier()
    # Steady-state loop as a prefetching training loop runs it: the plan phase of step s + 1 (plan_kernel: plans, resize
    # tables, labels -- a short latency-bound grid) is enqueued on a second stream and fills the tail of step s's big
    # kernel; events order plan(s) -> main(s) and main(s) -> plan(s + RING) (which rewrites that call's records / labels).
    # Every step's plan AND main phase runs inside the timed region (plan(0) is enqueued after ev0, main(K-1) before ev1).
    pipelined = os.environ.get("B200AUG_BENCH_SERIAL", "") == ""
    mode = {"two_streams": os.environ.get("B200AUG_BENCH_TWO_STREAMS", "") != ""}
    side = torch.cuda.Stream(dev)
    peer = torch.cuda.Stream(dev, priority=-1) if prio else side  # second stream of the overlapped-steps measurement (same priority)
    planned = [torch.cuda.Event() for _ in range(RING)]
    drained = [torch.cuda.Event() for _ in range(RING)]

    def timed_window():
        """EXACTLY args.steps steps between two events on the launching stream; returns milliseconds."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        if mode["two_streams"]:
            # consecutive steps alternate between two streams (each step: plan_kernel, then the big kernel behind a programmatic
            # dependent launch): steps are independent, so step s + 1 fills the SMs that step s's tail leaves idle
            peer.wait_event(ev0)
            for s in range(args.steps):
                calls[s % RING].launch((stream if s % 2 == 0 else peer).cuda_stream)  # (RING is even: a call stays on its stream)
            drained[0].record(peer)
            stream.wait_event(drained[0])
        elif pipelined:
            side.wait_event(ev0)
            calls[0].launch_plan(side.cuda_stream)
            planned[0].record(side)
            for s in range(args.steps):
                r, nxt = s % RING, (s + 1) % RING
                stream.wait_event(planned[r])
                calls[r].launch_main(stream.cuda_stream)
                drained[r].record(stream)
                if s + 1 < args.steps:
                    if s + 1 >= RING:
                        side.wait_event(drained[nxt])
                    calls[nxt].launch_plan(side.cuda_stream)
                    planned[nxt].record(side)
        else:
            for s in range(args.steps):
                calls[s % RING].launch(stream.cuda_stream)
        ev1.record(stream)
        barrier()
        return ev0.elapsed_time(ev1)

    # A window of K steps is ~0.1 ms x K: at the driver's K = 20 that is 2 ms, where one stray interrupt moves the number
    # by percent.  The K-step window is therefore timed `windows` times back to back (each with its own barriers and
    # events) and the MEDIAN window is reported; min / max and the per-rank medians go into the line.
    n_win = max(1, args.windows if args.windows > 0 else min(25, max(5, 2000 // max(args.steps, 1))))
    wins = [timed_window() for _ in range(n_win)]
    my_ms = float(np.median(wins))
    per_rank = [my_ms]
    if world > 1:
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = my_ms
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank = t.cpu().tolist()
        # the job's time for a window is the slowest rank's: max over ranks, window by window, then the median
        tw = torch.tensor(wins, device=dev, dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wins = tw.cpu().tolist()
    ms = float(np.median(wins))
    value = world * BATCH * args.steps / (ms * 1e-3)
    # for the record (single GPU): the same K steps with consecutive steps on alternating streams, so that step s + 1 fills the
    # SMs step s's tail leaves idle -- throughput of overlapped steps, NOT a per-launch time; it does not enter `value`
    overlapped_ms = None
    if world == 1 and not args.quick and not mode["two_streams"]:
        mode["two_streams"] = True
        overlapped_ms = float(np.median([timed_window() for _ in range(max(3, n_win // 2))]))
        mode["two_streams"] = False

    # ---- end-to-end through the public API from pinned host buffers
    pinned = []
    for h in hosts[:2]:
        pinned.append(Batch(Metadata((SRC, SRC), BATCH, "bench", None, dict(cats)), {k: torch.from_numpy(v).pin_memory() for k, v in h.items()}))
    label_keys = ("roi", "coord", "pose", "pt3d_68")
    host_out = {k: torch.empty_like(pinned[0][k]).pin_memory() for k in label_keys}
    frame_bytes = pinned[0]["image"].numel()
    label_bytes = sum(v.numel() * v.element_size() for k, v in pinned[0].items() if k != "image")
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    e2e_steps = min(args.steps, 40)

    host_outs = [host_out, {k: torch.empty_like(v).pin_memory() for k, v in host_out.items()}]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]

    def measure_e2e(zero_copy: bool, pipelined: bool, bands: bool = False, boxes: bool = False):
        """Public API from pinned host buffers: per step the host->device copy of the step's frames + labels, the fused
        launch, and the device->host read of the step's labels.  pipelined: two streams alternate, the host waits for step
        s - 1 after enqueuing step s (what a prefetching loader does), so the copy engine never idles; otherwise one stream,
        synchronised every step.  zero_copy: the frames stay in pinned host memory and the kernel reads them in place."""
        aug = FusedPoseAugmentation(OUT, rotation_aug_angle=30.0, roi_override="original", enable_image_aug=True, device=dev,
                                    zero_copy_frames=zero_copy, upload_row_bands=bands)
        aug.upload_boxes = boxes  # 2-D copies of the touched boxes, or whole row bands
        rows, host_ms = [], []

        def e2e_step(s):
            t_h = time.perf_counter()
            st = streams[s % 2] if pipelined else torch.cuda.current_stream(dev)
            with torch.cuda.stream(st):
                out = aug(pinned[s % 2])
                rows.append(aug.uploaded_bytes)
                for k in label_keys:
                    host_outs[s % 2][k].copy_(out[k], non_blocking=True)
            host_ms.append((time.perf_counter() - t_h) * 1e3)  # host work of the step: sampling, marshalling, enqueueing
            if pipelined:
                streams[(s - 1) % 2].synchronize()
            else:
                torch.cuda.synchronize(dev)

        for s in range(3):
            e2e_step(s)
        barrier()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            e2e_step(s)
        torch.cuda.synchronize(dev)
        mine = time.perf_counter() - t0
        barrier()
        e2e_s = time.perf_counter() - t0
        per_rank_s = [mine]
        if world > 1:
            t = torch.zeros(world + 1, device=dev, dtype=torch.float64)
            t[rank] = mine
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            per_rank_s = t[:world].cpu().tolist()
            t = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        return dict(value=world * BATCH * e2e_steps / e2e_s, rows=float(np.mean(rows[-e2e_steps:])), per_rank_s=per_rank_s,
                    host_ms=float(np.median(host_ms[-e2e_steps:])))

    e2e = measure_e2e(False, True, bands=True)
    e2e_value, band_rows = e2e["value"], e2e["rows"]
    h2d = label_bytes + int(band_rows)
    e2e_variants = {}
    if world == 1 and not args.quick:  # the other ways of getting the frames across, for the record (single GPU only)
        boxes_2d = measure_e2e(False, True, bands=True, boxes=True)
        e2e_variants = {
            "boxes_2d": {"value": boxes_2d["value"], "h2d_bytes_per_step": label_bytes + int(boxes_2d["rows"]),
                         "note": "same, only the columns of the touched boxes copied (b200aug_upload_boxes, one batched 2-D copy)"},
            "whole_frames": {"value": measure_e2e(False, True)["value"], "h2d_bytes_per_step": label_bytes + frame_bytes,
                             "note": "same, whole frames copied (upload_row_bands=False)"},
            "whole_frames_synchronised_every_step": {"value": measure_e2e(False, False)["value"]},
            "zero_copy_frames": {"value": measure_e2e(True, False)["value"],
                                 "h2d_bytes_per_step": label_bytes + int(alg_bytes - BATCH * (OUT * OUT * 4 + LABEL_BYTES)),
                                 "note": "frames left in pinned host memory and read in place by the kernel (only the view boxes "
                                         "cross PCIe), synchronised every step"}}
    clocks = sampler.stop() if sampler else None
    host_ms_all = [e2e["host_ms"]]
    if world > 1:
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = e2e["host_ms"]
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        host_ms_all = t.cpu().tolist()

    if rank == 0:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except (OSError, KeyError, ValueError):
            pass
        kernel_s = ms * 1e-3 / args.steps  # one fused kernel per step, event-timed on the launching stream
        achieved = alg_bytes / kernel_s / 1e9
        traffic = inst = None
        prof_src = None
        try:  # constants from the committed ncu capture of this kernel (profiles/), NOT measured in this run
            with open(os.path.join(ROOT, "profiles", "latest_traffic.json")) as f:
                prof = json.load(f)
            traffic, inst, prof_src = prof.get("dram_bytes_per_launch"), prof.get("warp_instructions_per_launch"), prof.get("source")
        except (OSError, ValueError):
            pass
        sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
        issue_peak = 148 * 4 * sm_hz  # warp instructions per second the 148 x 4 schedulers can issue
        line = {
            "metric": "augmented_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(world),
            "details": {"l2": f"inputs larger than L2: ring of {RING} distinct batches ({RING * BATCH * SRC * SRC / 1e6:.0f} MB of sources)",
                        "parallelism": f"per-sample sharding over {world} GPU(s), no collective; every rank works on the same "
                                       "four batches and draws (weak scaling: fixed work per GPU)",
                        "loop": ("plan phase of step s+1 on a second (lower-priority) stream next to the main phase of step s" if pipelined
                                 else "plan and main phase back to back on one stream"),
                        "timing": f"median of {len(wins)} windows of exactly {args.steps} steps (CUDA events on the launching stream, "
                                  "max over ranks per window)",
                        "window_ms": {"min": float(min(wins)), "median": ms, "max": float(max(wins))},
                        "per_rank_ms_per_step": [m / args.steps for m in per_rank],
                        **({"overlapped_steps": {"ms_per_step": overlapped_ms / args.steps, "samples_per_s": BATCH * args.steps / (overlapped_ms * 1e-3),
                                                 "note": "consecutive steps on two alternating streams (step s+1 fills the tail of step s); "
                                                         "throughput only, not the reported value"}} if overlapped_ms else {}),
                        "cpu_affinity_cores": affinity,
                        **({"EXPERIMENT_dropped": os.environ["B200AUG_BENCH_DROP"]} if os.environ.get("B200AUG_BENCH_DROP") else {})},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "note": "pinned host frames+labels -> FusedPoseAugmentation (host param sampling; host->device copy of the labels "
                            "and of the frame rows the sampled view boxes touch, b200aug_upload_row_bands; one fused launch) -> labels "
                            "read back into pinned host memory; two streams alternate and the host waits for step s-1 after enqueuing "
                            "step s, as a prefetching loader does",
                    "per_rank": {"samples_per_s": [BATCH * e2e_steps / t for t in e2e["per_rank_s"]],
                                 "h2d_GBps": [h2d * e2e_steps / t / 1e9 for t in e2e["per_rank_s"]],
                                 "host_ms_per_step": host_ms_all},
                    **e2e_variants},
            "gpu_launches": 2 * args.steps,  # plan_kernel + fused_augment_kernel per step
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": (f"{prof_src} (ncu capture of the same kernel and workload, a constant -- "
                                                                "not measured in this run)" if traffic else None),
                         "kernel": "fused_augment_kernel (+ plan_kernel on the side stream)", "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_us": kernel_s * 1e6, "peak_source": peak_src,
                         # what actually bounds the kernel: warp instructions issued against what the schedulers can issue
                         "issue": ({"warp_instructions_per_launch": inst, "peak_warp_inst_per_s": issue_peak,
                                    "achieved_warp_inst_per_s": inst / kernel_s, "frac": inst / kernel_s / issue_peak,
                                    "floor_us_at_full_issue": inst / issue_peak * 1e6, "source": prof_src} if inst else None)},
            "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="augment", choices=["augment", "train"],
                    help="augment: the hot path alone (the contract line); train: configs[2]/[3], ResNet-18 training step with the "
                         "augmentation on a side stream and DDP all-reduce (scripts/train_bench.py)")
    ap.add_argument("--rowbuf", type=int, default=0, help="row-buffer capacity per warp slot (bytes), 0 = library default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--spin-s", type=float, default=0.6, help="minimum seconds of untimed warm-up launches (clock settling)")
    ap.add_argument("--windows", type=int, default=0, help="how many times the K-step window is timed (median reported); 0 = auto")
    ap.add_argument("--no-pin", action="store_true", help="do not bind ranks to their GPU's host cores (N > 1)")
    ap.add_argument("--quick", action="store_true", help="skip the extra single-GPU e2e variants")
    args = ap.parse_args()
    if args.workload == "train":
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import train_bench

        world = int(os.environ.get("WORLD_SIZE", "1"))
        line = train_bench.run(256 if world > 1 else 128, max(args.steps, 10) if args.steps != 200 else 40, max(args.warmup, 5), False)
        if line is not None:
            print(json.dumps(line))
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
