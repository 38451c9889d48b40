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
