"""Device-side synthetic visibility cubes for benchmarks (value distribution of the
reference generator's 1024 x 1024 recipe, synthetic_generator.py:520-656 and
configs/data_generation/synthetic_val_1k.yaml:9-25; distribution parity only -- the host
MT19937 stream of the reference cannot be reproduced on a device).

Per baseline: amplitude N(1, 0.1) x 8th-order bandpass with 10 % edges (edge rows exactly
zero), 20 narrow-band persistent, 5 broad-band persistent, 20 narrow-band bursty and 5
broad-band bursty rectangles plus one linear sweep at U(1e6, 1e7) mJy; pol 0 full RFI,
pol 1 0.8 x RFI, pols 2.. noise only; phase U(0, 2 pi); complex64.  Philox generator,
seed = `seed`, one subsequence per baseline.
"""
from __future__ import annotations

import math

import torch


def _bandpass(channels, device, order=8, edge_fraction=0.1):
    bp = torch.ones(channels, device=device, dtype=torch.float32)
    edge = int(channels * edge_fraction)
    if edge > 0:
        t = torch.arange(edge, device=device, dtype=torch.float32) / edge
        bp[:edge] = t**order
        bp[channels - edge:] = torch.flip(t**order, dims=[0])
    return bp


def device_cube(n_bl, n_pol, channels, times, seed=1234, device="cuda", rfi=True):
    """-> (complex64 cube (n_bl, n_pol, C, T), bool mask of the same shape), both on `device`."""
    device = torch.device(device)
    cube = torch.empty((n_bl, n_pol, channels, times), dtype=torch.complex64, device=device)
    mask = torch.zeros((n_bl, n_pol, channels, times), dtype=torch.bool, device=device)
    bp = _bandpass(channels, device)
    host = torch.Generator(device="cpu")
    for b in range(n_bl):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + b)
        host.manual_seed(seed * 1000003 + b)
        base = (1.0 + 0.1 * torch.randn((channels, times), generator=g, device=device)) * bp[:, None]
        sig = torch.zeros((channels, times), device=device)
        m = torch.zeros((channels, times), dtype=torch.bool, device=device)
        if rfi:
            def u(lo, hi):
                return float(torch.empty(1).uniform_(lo, hi, generator=host))

            def ri(lo, hi):
                return int(torch.randint(lo, hi, (1,), generator=host))

            c_lo, c_hi = int(channels * 0.1), int(channels * 0.9)
            t_lo, t_hi = int(times * 0.1), int(times * 0.9)
            for _ in range(20):  # narrow-band persistent
                c, w = ri(c_lo, c_hi), ri(1, 10)
                sl = slice(max(0, c - w // 2), min(channels, c + w // 2 + 1))
                sig[sl, :] += u(1e6, 1e7); m[sl, :] = True
            for _ in range(5):  # broad-band persistent
                c, w = ri(t_lo, t_hi), ri(5, 50)
                sl = slice(max(0, c - w // 2), min(times, c + w // 2))
                sig[:, sl] += u(1e6, 1e7); m[:, sl] = True
            for _ in range(20):  # narrow-band bursty
                c, w = ri(c_lo, c_hi), ri(2, 20)
                fs = slice(max(0, c - w // 2), min(channels, c + w // 2))
                amp = u(1e6, 1e7)
                for _ in range(ri(3, 15)):
                    t0, d = ri(0, times), ri(2, 20)
                    ts = slice(max(0, t0 - d // 2), min(times, t0 + d // 2))
                    sig[fs, ts] += amp; m[fs, ts] = True
            for _ in range(5):  # broad-band bursty
                amp = u(1e6, 1e7)
                for _ in range(ri(2, 10)):
                    t0, d = ri(0, times), ri(1, 5)
                    ts = slice(max(0, t0 - d // 2), min(times, t0 + d // 2))
                    sig[:, ts] += amp; m[:, ts] = True
            # one linear frequency sweep
            f0, f1, w = ri(c_lo, channels // 2), ri(channels // 2, c_hi), ri(2, 10)
            tt = torch.arange(times, device=device)
            centre = (f0 + (f1 - f0) * tt.float() / times).long()
            rows = torch.arange(channels, device=device)[:, None]
            sweep = (rows >= (centre - w // 2)[None, :]) & (rows < (centre + w // 2)[None, :])
            sig = sig + sweep.float() * u(1e6, 1e7)
            m |= sweep
        for p in range(n_pol):
            if p == 0:
                real, mk = base + sig, m
            elif p == 1:
                real = 0.8 * sig + 0.2 * 0.1 * torch.randn(sig.shape, generator=g, device=device) + base
                mk = m
            else:
                real = 1.0 + 0.1 * torch.randn(sig.shape, generator=g, device=device)
                mk = torch.zeros_like(m)
            phase = torch.rand(sig.shape, generator=g, device=device) * (2 * math.pi)
            cube[b, p] = torch.polar(real.abs(), phase)
            mask[b, p] = mk
    return cube, mask
