"""N > 1 on real GPUs (needs >= 2 devices; skipped on a single-GPU box): the metric counts summed
over baseline shards by the fused count + all-reduce kernel (NVLink peer memory,
`rfi_confusion_counts_allreduce`) against the counts of the unsharded masks and against the NCCL
path.  One process per GPU, torch.distributed over NCCL for the plumbing."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from rfi_toolbox_b200.evaluation.metrics import _counts_tensor, confusion_counts, evaluate_segmentation
    from rfi_toolbox_b200.utils.sharding import baseline_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    out = []
    try:
        for epoch, n_bl in enumerate([5, 5, 4, 1, 7, 5, 5]):      # n_bl = 1: rank 1 holds an EMPTY shard
            g = torch.Generator().manual_seed(100 + epoch)
            true = (torch.rand((n_bl, 2, 128, 256), generator=g) < 0.1).to(torch.uint8)
            pred = true ^ (torch.rand(true.shape, generator=g) < 0.03).to(torch.uint8)
            sl = baseline_shard(n_bl, world, rank)
            want = tuple(_counts_tensor(pred.to(dev), true.to(dev)).tolist())
            got = confusion_counts(pred[sl].to(dev), true[sl].to(dev), group=True)
            os.environ["RFI_NO_PEER"] = "1"
            nccl = confusion_counts(pred[sl].to(dev), true[sl].to(dev), group=True)
            del os.environ["RFI_NO_PEER"]
            m = evaluate_segmentation(pred[sl].to(dev), true[sl].to(dev), group=True)
            out.append((want, tuple(got), tuple(nccl), float(m["iou"])))
        from rfi_toolbox_b200.utils.peer import _CONTEXTS
        fused = all(c is not None for c in _CONTEXTS.values()) and len(_CONTEXTS) > 0
        q.put((rank, out, fused))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_count_allreduce_matches_global_counts(native_lib):
    import torch.multiprocessing as mp
    world, port = 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, out, fused in results:
        assert fused, "the peer-memory exchange was not used (fell back to NCCL)"
        for want, got, nccl, iou in out:
            assert got == want and nccl == want
            tp, fp, fn = want
            assert np.isclose(iou, tp / (tp + fp + fn))


def _worker_partial_failure(rank, world, port, q):
    """One rank cannot set up its exchange buffer: EVERY rank must fall back to NCCL (no hang)."""
    import torch.distributed as dist

    from rfi_toolbox_b200.evaluation.metrics import _counts_tensor, confusion_counts
    from rfi_toolbox_b200.utils.sharding import baseline_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RFI_PEER_FAIL_RANK="1")
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(5)
        true = (torch.rand((6, 2, 128, 128), generator=g) < 0.1).to(torch.uint8)
        pred = true ^ (torch.rand(true.shape, generator=g) < 0.03).to(torch.uint8)
        sl = baseline_shard(6, world, rank)
        want = tuple(_counts_tensor(pred.to(dev), true.to(dev)).tolist())
        got = [tuple(confusion_counts(pred[sl].to(dev), true[sl].to(dev), group=True)) for _ in range(3)]
        from rfi_toolbox_b200.utils.peer import _CONTEXTS
        q.put((rank, want, got, [c is None for c in _CONTEXTS.values()]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partial_peer_failure_falls_back_everywhere(native_lib):
    import torch.multiprocessing as mp
    world, port = 2, 29633
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_partial_failure, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, want, got, none_ctx in results:
        assert none_ctx == [True], "a rank kept the peer path although rank 1 failed"
        assert all(g == want for g in got)


def _worker_sharded_ffi(rank, world, port, q):
    """compute_ffi / compute_statistics of a cube sharded by baseline == the unsharded oracle."""
    import torch.distributed as dist

    import oracle
    from rfi_toolbox_b200 import compute_ffi, compute_statistics
    from rfi_toolbox_b200.utils.sharding import baseline_shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    out = []
    try:
        for case, (n_bl, dtype) in enumerate([(5, np.complex64), (4, np.float32), (1, np.complex64)]):   # n_bl = 1: an empty shard
            rng = np.random.default_rng(40 + case)
            shape = (n_bl, 2, 96, 130)
            data = rng.normal(0, 1, shape) + 1j * rng.normal(0, 1, shape)
            flags = rng.random(shape) < 0.1
            data[flags] *= 30.0
            flags ^= rng.random(shape) < 0.01
            data = (data if np.dtype(dtype).kind == "c" else np.abs(data)).astype(dtype)
            sl = baseline_shard(n_bl, world, rank)
            d, f = torch.from_numpy(data[sl]).to(dev), torch.from_numpy(flags[sl]).to(dev)
            got_ffi = compute_ffi(d, f, group=True)
            got_b = compute_statistics(d, None, group=True)
            got_a = compute_statistics(d, f, group=True)
            out.append((got_ffi, got_b, got_a, oracle.compute_ffi(data, flags), oracle.compute_statistics(data, None),
                        oracle.compute_statistics(data, flags)))
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _check_sharded(out):
    for got_ffi, got_b, got_a, want_ffi, want_b, want_a in out:
        for got, want in ((got_b, want_b), (got_a, want_a)):
            for k in ("median", "mad"):
                assert np.float32(got[k]) == np.float32(want[k]), (k, got[k], want[k])
            for k in ("mean", "std", "flagged_fraction"):
                assert got[k] == pytest.approx(want[k], rel=1e-6, abs=1e-7), (k, got[k], want[k])
            assert got["count"] == want["count"]
        for k, v in want_ffi.items():
            assert got_ffi[k] == pytest.approx(v, rel=1e-5, abs=1e-7), (k, got_ffi[k], v)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_ffi_matches_unsharded_oracle(native_lib):
    import torch.multiprocessing as mp
    world, port = 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded_ffi, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, out in results:
        _check_sharded(out)
    assert results[0][1][0][0] == results[1][1][0][0]     # every rank holds the same FFI


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_sharded_ffi_group_of_one(native_lib):
    """The sharded code path (radix select with all-reduced counts) on a process group of ONE rank: runs on the
    single-GPU box of the regular GPU suite."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker_sharded_ffi, args=(0, 1, 29643, q))
    p.start()
    rank, out = q.get(timeout=300)
    p.join(timeout=120)
    assert p.exitcode == 0
    _check_sharded(out)
