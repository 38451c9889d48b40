"""N > 1 host logic on CPU: world_size-2 gloo run of the baseline sharding + count reduction
that bench.py / evaluate_segmentation(group=...) use over NCCL on the GPU box."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from rfi_toolbox_b200.utils.sharding import allreduce_counts, baseline_shard
from tests.cubes import make_cube


def test_baseline_shard_partition():
    for n in (1, 7, 45, 351):
        for world in (1, 2, 4, 8):
            parts = [baseline_shard(n, world, r) for r in range(world)]
            covered = [i for s in parts for i in range(s.start, s.stop)]
            assert covered == list(range(n))
            sizes = [s.stop - s.start for s in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [s.stop - s.start for s in (baseline_shard(351, 8, r) for r in range(8))] == [44] * 7 + [43]
    with pytest.raises(ValueError):
        baseline_shard(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, mask = make_cube(n_bl=5, n_pol=2, channels=128, times=128, seed=77)
    pred = mask ^ (np.random.default_rng(5).random(mask.shape) < 0.05)
    sl = baseline_shard(mask.shape[0], world, rank)
    tp, fp, fn = oracle.confusion_counts(pred[sl], mask[sl])
    counts = allreduce_counts(torch.tensor([tp, fp, fn], dtype=torch.int64))
    q.put((rank, counts.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_counts_sum_to_global_counts():
    world, port = 2, 29613
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, mask = make_cube(n_bl=5, n_pol=2, channels=128, times=128, seed=77)
    pred = mask ^ (np.random.default_rng(5).random(mask.shape) < 0.05)
    want = list(oracle.confusion_counts(pred, mask))
    assert all(c == want for _, c in results)
