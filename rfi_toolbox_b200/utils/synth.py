"""Bench / script convenience over `rfi_toolbox_b200.data_generation.SyntheticDataGenerator`
(the hand-written device generator, csrc/rfi_synth.cu): `device_cube` keeps its old signature."""
from __future__ import annotations


def device_cube(n_bl, n_pol, channels, times, seed=1234, device="cuda", rfi=True, first_baseline=0):
    """-> (complex64 cube (n_bl, n_pol, C, T), bool mask of the same shape), both on `device`.
    Baseline b of the cube depends on (seed, first_baseline + b) only (sharding invariant)."""
    from ..data_generation import SyntheticDataGenerator
    cube, mask, _ = SyntheticDataGenerator(device=device).generate_cube(
        n_bl, channels, times, num_polarizations=n_pol, seed=seed, first_baseline=first_baseline, rfi=rfi)
    return cube, mask
