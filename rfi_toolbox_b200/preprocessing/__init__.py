"""Preprocessing operators (mirror of rfi_toolbox/preprocessing/__init__.py:7)."""
from .preprocessor import Preprocessor, canonical_index_map, iter_dataset_chunks, patchify

__all__ = ["Preprocessor", "patchify", "canonical_index_map", "iter_dataset_chunks"]
