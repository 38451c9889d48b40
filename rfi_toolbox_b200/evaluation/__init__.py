"""Evaluation operators (mirror of rfi_toolbox/evaluation/__init__.py:8-21)."""
from .metrics import (
    compute_dice,
    compute_f1,
    compute_iou,
    compute_precision,
    compute_recall,
    confusion_counts,
    confusion_counts_async,
    evaluate_segmentation,
    evaluate_segmentation_async,
    evaluate_segmentation_batch,
)
from .statistics import (
    compute_calcquality,
    compute_ffi,
    compute_ffi_batch,
    compute_mad,
    compute_statistics,
    compute_statistics_batch,
    evaluate_pairs,
    print_statistics_comparison,
)

__all__ = [
    "compute_iou", "compute_precision", "compute_recall", "compute_f1", "compute_dice",
    "evaluate_segmentation", "evaluate_segmentation_async", "evaluate_segmentation_batch", "confusion_counts", "confusion_counts_async",
    "compute_statistics", "compute_ffi", "compute_mad", "compute_statistics_batch", "compute_ffi_batch", "compute_calcquality", "print_statistics_comparison", "evaluate_pairs",
]
