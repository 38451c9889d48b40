"""Oracle (test infrastructure): NumPy restatement of evaluation/metrics.py:18-172.

`confusion_counts` is the integer core the CUDA kernel is diffed against (bit-exact);
`metrics_from_counts` applies the reference's float64 formulas and guard branches."""

from __future__ import annotations

import numpy as np


def _as_bool(a):
    """metrics.py:18-22 + `.astype(bool)` (:36-37): any non-zero (NaN included) is True."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a).astype(bool)


def confusion_counts(pred, true):
    """TP, FP, FN as Python ints (metrics.py:66-68)."""
    p, t = _as_bool(pred), _as_bool(true)
    tp = int(np.logical_and(p, t).sum())
    fp = int(np.logical_and(p, ~t).sum())
    fn = int(np.logical_and(~p, t).sum())
    return tp, fp, fn


def metrics_from_counts(tp, fp, fn):
    """The five ratios with the reference's guards and return types
    (np.float64 from the divisions, Python float from the guard branches)."""
    tp, fp, fn = np.int64(tp), np.int64(fp), np.int64(fn)
    union = tp + fp + fn
    iou = 1.0 if union == 0 else tp / union  # metrics.py:39-45
    if tp + fp == 0:  # metrics.py:70-79
        precision = 1.0 if fn == 0 else 0.0
    else:
        precision = tp / (tp + fp)
    recall = 1.0 if tp + fn == 0 else tp / (tp + fn)  # metrics.py:101-104
    if precision + recall == 0:  # metrics.py:120-126
        f1 = 0.0
    else:
        f1 = 2 * (precision * recall) / (precision + recall)
    dice = 1.0 if 2 * tp + fp + fn == 0 else (2 * tp) / (2 * tp + fp + fn)  # metrics.py:149-152
    return {"iou": iou, "precision": precision, "recall": recall, "f1": f1, "dice": dice}


def evaluate_segmentation(pred, true):
    """metrics.py:155-172."""
    return metrics_from_counts(*confusion_counts(pred, true))
