"""Seeded synthetic cubes shared by the tests, the smoke test and bench.py.

`make_cube` follows the value distribution of the reference generator's 1024 x 1024 recipe
(rfi_toolbox/data_generation/synthetic_generator.py:520-656, configs/.../synthetic_val_1k.yaml):
N(1, 0.1) noise x 8th-order bandpass with exact-zero edge rows, narrow/broad-band persistent
and bursty rectangles at 1e6..1e7 mJy, pol 1 = 0.8 x RFI, pols 2.. noise only, uniform phase.
It is NOT the reference generator (which draws from the global MT19937 stream); parity tests
only need both sides to see the SAME array.
"""
from __future__ import annotations

import numpy as np


def make_cube(n_bl=2, n_pol=2, channels=256, times=384, seed=0, dtype=np.complex64, rfi=True,
              special=False):
    rng = np.random.default_rng(seed)
    out = np.empty((n_bl, n_pol, channels, times), dtype=np.complex128)
    mask = np.zeros((n_bl, n_pol, channels, times), dtype=bool)
    edge = max(int(channels * 0.1), 1)
    bp = np.ones(channels)
    t = np.arange(edge) / edge
    bp[:edge] = t**8
    bp[::-1][:edge] = t**8
    for b in range(n_bl):
        base = rng.normal(1.0, 0.1, (channels, times)) * bp[:, None]
        sig = np.zeros((channels, times))
        m = np.zeros((channels, times), dtype=bool)
        if rfi:
            for _ in range(6):  # narrow-band persistent
                c = rng.integers(int(channels * 0.1), int(channels * 0.9)); w = rng.integers(1, 6)
                sig[c:c + w, :] += rng.uniform(1e6, 1e7); m[c:c + w, :] = True
            for _ in range(3):  # broad-band persistent
                c = rng.integers(int(times * 0.1), int(times * 0.9)); w = rng.integers(2, 20)
                sig[:, c:c + w] += rng.uniform(1e6, 1e7); m[:, c:c + w] = True
            for _ in range(8):  # bursts
                c = rng.integers(0, channels - 8); w = rng.integers(2, 8)
                t0 = rng.integers(0, times - 16); d = rng.integers(2, 16)
                sig[c:c + w, t0:t0 + d] += rng.uniform(1e6, 1e7); m[c:c + w, t0:t0 + d] = True
        for p in range(n_pol):
            if p == 0:
                real, mk = base + sig, m
            elif p == 1:
                real, mk = 0.8 * sig + 0.2 * rng.normal(0, 0.1, sig.shape) + base, m
            else:
                real, mk = rng.normal(1.0, 0.1, sig.shape), np.zeros_like(m)
            out[b, p] = real * np.exp(1j * rng.uniform(0, 2 * np.pi, real.shape))
            mask[b, p] = mk
    if special:  # NaN / inf / signed zero samples
        out[0, 0, 5, 7] = np.nan
        out[0, 0, 200 % channels, 300 % times] = np.inf
        out[-1, -1, 130 % channels, 5] = 0.0
    dt = np.dtype(dtype)
    if dt.kind == "c":
        return out.astype(dt), mask
    return np.abs(out).astype(dt), mask
