"""rfi_toolbox_b200 -- B200-native (sm_100a) drop-in for the rfi_toolbox preprocessing and
evaluation hot path: `Preprocessor(data, flags).create_dataset(...)`,
`evaluate_segmentation`, `compute_ffi` / `compute_statistics`.

Host side is Python (as in the reference); the arithmetic runs in hand-written CUDA kernels
behind the C ABI of `include/rfi_b200.h` (`_lib/librfi_b200.so`, bound with ctypes).
No CPU fallback, no Triton, no multi-backend dispatch.
"""

__version__ = "0.1.0"

from . import data_generation, datasets, evaluation, preprocessing  # noqa: F401,E402
from .datasets import TorchDataset  # noqa: F401,E402
from .evaluation import (  # noqa: F401,E402
    compute_dice,
    compute_f1,
    compute_ffi,
    compute_ffi_batch,
    compute_iou,
    compute_precision,
    compute_recall,
    compute_statistics,
    compute_statistics_batch,
    evaluate_pairs,
    evaluate_segmentation,
    evaluate_segmentation_async,
    evaluate_segmentation_batch,
)
from .data_generation import SyntheticDataGenerator  # noqa: F401,E402
from .preprocessing import GPUPreprocessor, Preprocessor, iter_dataset_chunks  # noqa: F401,E402
