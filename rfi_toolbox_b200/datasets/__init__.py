"""Dataset containers (mirror of rfi_toolbox/datasets/__init__.py:7)."""
from .batched_dataset import BatchWriter, TorchDataset

__all__ = ["TorchDataset", "BatchWriter"]
