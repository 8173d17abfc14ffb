"""world_size-2 `gloo` run of the N>1 host logic: per-sample sharding with no data-path collective.

Each rank draws the parameters of its own slice of a global batch and runs the CPU oracle on it (the oracle stands in for
the kernel here -- this is the checker's side of the fence); rank 0 gathers and checks that the sharded result equals the
single-process result bit for bit: the shards are disjoint and cover the batch, and the Philox noise of a sample is keyed
by its GLOBAL id, so it does not depend on the world size.  Also covers bench.py's max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "neuralnet-tracker-traincode_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import photometric as opho
        from trackertraincode_b200.datatransformation import sharding

        assert sharding.rank_world() == (rank, world)
        n_global, step, local = 6, 3, 3
        lo, hi = sharding.shard_range(n_global, rank, world)
        assert hi - lo == local
        # every rank can derive the global ids of its samples without talking to anyone
        off = sharding.global_sample_offset(step, rank, world, local)
        rng = np.random.default_rng(11)  # same global draw on every rank, each keeps its slice
        x = rng.random((n_global, 1, 16, 16)).astype(np.float32)
        pp = opho.sample_photo_params(rng, n_global, seed=99, sample_offset=sharding.global_sample_offset(step, 0, world, local))
        pp.noise_apply[:, 0] = True
        mine = opho.photometric_batch(x[lo:hi], pp.slice(lo, hi))
        assert pp.slice(lo, hi).sample_offset == off
        gathered = [torch.zeros(local, 1, 16, 16) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))
        t = sharding.max_over_ranks(1.0 + rank)  # slowest rank defines the step time
        ids = [None] * world
        dist.all_gather_object(ids, list(range(off, off + local)))
        if rank == 0:
            whole = opho.photometric_batch(x, pp)
            q.put(dict(equal=bool(np.array_equal(torch.cat(gathered).numpy(), whole)), t=t, ids=sorted(sum(ids, []))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_augmentation_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    res = q.get()
    assert res["equal"], "sharded result differs from the single-process result"
    assert res["t"] == 2.0
    step, local = 3, 3
    assert res["ids"] == list(range(step * world * local, (step + 1) * world * local))  # disjoint, contiguous global ids
