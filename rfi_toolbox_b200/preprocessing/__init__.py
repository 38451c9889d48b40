"""Preprocessing operators (mirror of rfi_toolbox/preprocessing/__init__.py:7)."""
from .gpu_preprocessor import GPUPreprocessor
from .preprocessor import Preprocessor, canonical_index_map, iter_dataset_chunks, patchify

__all__ = ["Preprocessor", "GPUPreprocessor", "patchify", "canonical_index_map", "iter_dataset_chunks"]
