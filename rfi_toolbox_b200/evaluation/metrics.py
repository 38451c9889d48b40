"""Segmentation metrics -- drop-in for rfi_toolbox/evaluation/metrics.py:25-172.

Same names, arguments and return values (`np.float64` from the divisions, Python `float`
from the guard branches).  The boolean reductions run on the GPU in one pass
(`rfi_confusion_counts`, csrc/rfi_metrics.cu); the five ratios are then formed on the host
in float64 with the reference's own guard branches, which reproduces it to the last bit
(including f1 != dice in the last ulp, metrics.py:120-126 vs :145-152).

Extensions (not in the reference): `confusion_counts`, `evaluate_segmentation_async`, `evaluate_segmentation_batch`
(per-pair sweep in one launch) and the `group=` keyword, which all-reduces {TP, FP, FN}
over a torch.distributed process group when masks are sharded by baseline across GPUs.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..utils.device import as_device_tensor, current_stream_ptr, require_cuda

_FLOAT_DTYPES = {torch.float16, torch.bfloat16, torch.float32, torch.float64}
_INT_DTYPES = {torch.bool, torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64}


def _mask_operand(x, device):
    """-> (contiguous device tensor, element size, is_float) with astype(bool) semantics."""
    if isinstance(x, torch.Tensor) and x.is_cuda:
        device = x.device
    t = as_device_tensor(x, device)
    if t.dtype in _FLOAT_DTYPES:
        return t, t.element_size(), 1
    if t.dtype in _INT_DTYPES or t.dtype in (torch.uint16, torch.uint32, torch.uint64):
        return t, t.element_size(), 0
    raise TypeError(f"unsupported mask dtype {t.dtype}")


def _pick_device(pred, true):
    for x in (pred, true):
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return require_cuda()


def _counts_tensor(pred, true, n_seg=None, seg=None):
    lib = _native.load()
    device = _pick_device(pred, true)
    require_cuda(device)
    p, ep, fp = _mask_operand(pred, device)
    t, et, ft = _mask_operand(true, device)
    if p.numel() != t.numel():
        # NumPy would broadcast or raise; the reference is only ever called with equal shapes
        raise ValueError(f"pred and true must have the same number of elements ({p.numel()} vs {t.numel()})")
    with torch.cuda.device(device):
        stream = current_stream_ptr(device)
        if n_seg is None:
            counts = torch.zeros(3, dtype=torch.int64, device=device)
            rc = lib.rfi_confusion_counts(p.data_ptr(), ep, fp, t.data_ptr(), et, ft, p.numel(),
                                          counts.data_ptr(), stream)
            _native.check(rc, "rfi_confusion_counts")
        else:
            counts = torch.empty((n_seg, 3), dtype=torch.int64, device=device)
            rc = lib.rfi_confusion_counts_segmented(p.data_ptr(), ep, fp, t.data_ptr(), et, ft,
                                                    n_seg, seg, counts.data_ptr(), stream)
            _native.check(rc, "rfi_confusion_counts_segmented")
    return counts


def _counts_allreduced(pred, true, ctx):
    """{TP, FP, FN, missing} (device int64[4]) summed over the ranks of `ctx` by ONE kernel: the
    reduction's last CTA exchanges the totals through NVLink peer memory
    (`rfi_confusion_counts_allreduce`)."""
    lib = _native.load()
    device = ctx.device
    p, ep, fp = _mask_operand(pred, device)
    t, et, ft = _mask_operand(true, device)
    if p.numel() != t.numel():
        raise ValueError(f"pred and true must have the same number of elements ({p.numel()} vs {t.numel()})")
    with torch.cuda.device(device):
        counts = torch.empty(4, dtype=torch.int64, device=device)
        rc = lib.rfi_confusion_counts_allreduce(p.data_ptr(), ep, fp, t.data_ptr(), et, ft, p.numel(),
                                                ctx.ptrs, ctx.world, ctx.rank, ctx.next_epoch(),
                                                counts.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_confusion_counts_allreduce")
    return counts, ctx.epoch


class PendingCounts:
    """{TP, FP, FN} of a `confusion_counts_async` call: the counting kernel and the copy of its three
    totals to pinned host memory are enqueued; `result()` waits for that copy only -- not for
    whatever the caller enqueued afterwards (a plain `.tolist()` would drain the stream)."""

    def __init__(self, counts_dev, epoch=None):
        self._epoch = epoch
        self._host = torch.empty(counts_dev.shape, dtype=torch.int64, pin_memory=True)
        self._host.copy_(counts_dev, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()
        self._value = None

    def result(self):
        if self._value is None:
            self._event.synchronize()
            vals = self._host.tolist()
            if len(vals) == 4 and vals[3]:
                raise RuntimeError(f"{vals[3]} rank(s) never reached the metric exchange (epoch {self._epoch})")
            self._value = tuple(vals[:3])
            self._host = None
        return self._value


class PendingMetrics:
    """`evaluate_segmentation_async` handle; `result()` -> the metric dict."""

    def __init__(self, counts):
        self._counts = counts

    def result(self):
        return _ratios(*self._counts.result())


_METRIC_STREAMS = {}


def _metrics_stream(device):
    s = _METRIC_STREAMS.get(device.index)
    if s is None:
        s = _METRIC_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return s


def confusion_counts_async(pred, true, group=None, side_stream=False):
    """`confusion_counts` without the host synchronisation: -> `PendingCounts`.  With `group`, the
    sum over the ranks happens inside the counting kernel (NVLink peer memory); `RFI_NO_PEER=1`,
    several hosts or missing CUDA IPC fall back to an NCCL all-reduce enqueued behind the kernel.

    `side_stream=True`: the counting kernel (HBM-bound, a few registers, no shared memory) and its
    24-byte download run on a per-device side stream that waits for the work queued so far, so that
    whatever the caller enqueues next on the current stream -- in a streaming loop the issue-bound
    statistics kernel of the next `create_dataset_async` -- shares the SMs with it."""
    if side_stream:
        device = _pick_device(pred, true)
        require_cuda(device)
        cur, side = torch.cuda.current_stream(device), _metrics_stream(device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            out = confusion_counts_async(pred, true, group=group)
        for x in (pred, true):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                x.record_stream(side)
        return out
    if group is not None:
        import os

        import torch.distributed as dist
        g = None if group is True else group
        if dist.get_world_size(g) > 1:
            device = _pick_device(pred, true)
            ctx = None
            if not os.environ.get("RFI_NO_PEER"):
                from ..utils.peer import peer_context
                ctx = peer_context(g, device)
            if ctx is not None:
                counts, epoch = _counts_allreduced(pred, true, ctx)
                with torch.cuda.device(device):
                    return PendingCounts(counts, epoch)
            counts = _counts_tensor(pred, true)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=g)
            with torch.cuda.device(counts.device):
                return PendingCounts(counts)
    counts = _counts_tensor(pred, true)
    with torch.cuda.device(counts.device):
        return PendingCounts(counts)


def confusion_counts(pred, true, group=None):
    """(TP, FP, FN) as Python ints; summed over `group` if given (`True` = the default process
    group).  On one box the sum happens inside the counting kernel, over NVLink peer memory;
    `RFI_NO_PEER=1`, several hosts or missing CUDA IPC fall back to an NCCL all-reduce."""
    return confusion_counts_async(pred, true, group=group).result()


def _ratios(tp, fp, fn):
    """metrics.py:39-45, 66-79, 98-104, 120-126, 145-152 on integer counts."""
    tp, fp, fn = np.int64(tp), np.int64(fp), np.int64(fn)
    union = tp + fp + fn
    iou = 1.0 if union == 0 else tp / union
    if tp + fp == 0:
        precision = 1.0 if fn == 0 else 0.0
    else:
        precision = tp / (tp + fp)
    recall = 1.0 if tp + fn == 0 else tp / (tp + fn)
    f1 = 0.0 if precision + recall == 0 else 2 * (precision * recall) / (precision + recall)
    dice = 1.0 if 2 * tp + fp + fn == 0 else (2 * tp) / (2 * tp + fp + fn)
    return {"iou": iou, "precision": precision, "recall": recall, "f1": f1, "dice": dice}


def evaluate_segmentation(pred, true, group=None):
    """metrics.py:155-172 -> {'iou','precision','recall','f1','dice'}."""
    return _ratios(*confusion_counts(pred, true, group=group))


def evaluate_segmentation_async(pred, true, group=None, side_stream=False):
    """`evaluate_segmentation` for streaming callers: the counting kernel and the download of its
    three totals are enqueued, the ratios are formed in `result()`.  Reading the metrics of step k
    after step k + 1 has been enqueued keeps the GPU queue from draining between steps."""
    return PendingMetrics(confusion_counts_async(pred, true, group=group, side_stream=side_stream))


def compute_iou(pred, true):
    return evaluate_segmentation(pred, true)["iou"]


def compute_precision(pred, true):
    return evaluate_segmentation(pred, true)["precision"]


def compute_recall(pred, true):
    return evaluate_segmentation(pred, true)["recall"]


def compute_f1(pred, true):
    return evaluate_segmentation(pred, true)["f1"]


def compute_dice(pred, true):
    return evaluate_segmentation(pred, true)["dice"]


def evaluate_segmentation_batch(pred, true):
    """Per-pair metrics for stacks of masks (N, ...): one launch, one CTA per pair.

    Returns a dict of float64 arrays of length N whose i-th entries equal
    `evaluate_segmentation(pred[i], true[i])` of the reference."""
    n = len(pred)
    if n == 0:
        return {k: np.zeros(0) for k in ("iou", "precision", "recall", "f1", "dice")}
    numel = pred.numel() if isinstance(pred, torch.Tensor) else np.asarray(pred).size
    seg = numel // n
    c = _counts_tensor(pred, true, n_seg=n, seg=seg).cpu().numpy()
    return _ratio_arrays(c[:, 0].astype(np.int64), c[:, 1].astype(np.int64), c[:, 2].astype(np.int64))


def _ratio_arrays(tp, fp, fn):
    """`_ratios` over int64 arrays of counts (same branches, same float64 operations)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        union = tp + fp + fn
        iou = np.where(union == 0, 1.0, tp / union)
        precision = np.where(tp + fp == 0, np.where(fn == 0, 1.0, 0.0), tp / (tp + fp))
        recall = np.where(tp + fn == 0, 1.0, tp / (tp + fn))
        f1 = np.where(precision + recall == 0, 0.0, 2 * (precision * recall) / (precision + recall))
        dice = np.where(2 * tp + fp + fn == 0, 1.0, (2 * tp) / (2 * tp + fp + fn))
    return {"iou": iou, "precision": precision, "recall": recall, "f1": f1, "dice": dice,
            "tp": tp, "fp": fp, "fn": fn}
