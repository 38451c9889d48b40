"""Synthetic data generation (mirror of rfi_toolbox/data_generation/__init__.py)."""
from .synthetic_generator import DEFAULT_RFI_COUNTS, SyntheticDataGenerator, draw_rfi_events

__all__ = ["SyntheticDataGenerator", "draw_rfi_events", "DEFAULT_RFI_COUNTS"]
