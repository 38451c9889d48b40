"""SyntheticDataGenerator -- device-side synthetic visibility cubes, the step before the hot path
(rfi_toolbox/data_generation/synthetic_generator.py:520-815).

What is kept from the reference: the recipe of `_generate_single_sample` (:520-656) -- clean
amplitude N(noise, 0.1 noise) x polynomial bandpass with exact-zero edge rows, the six RFI event
types with the parameter ranges of `_add_*` (:675-815) and the counts of
`configs/data_generation/synthetic_val_1k.yaml:15-20`, amplitudes U(rfi_power_min, rfi_power_max)
x 1000 mJy, pol 0 full RFI, pol 1 `pol_corr` x RFI, pols >= 2 noise only, uniform phase -- and the
return shapes `(B, Npol, C, T)` complex + exact bool mask + the list of event parameters.

What changes: the pixels are drawn on the GPU (`rfi_synth_waterfalls`, one Philox4x32-10 counter
per pixel) straight into HBM, complex64, many baselines per launch; only the few hundred event
parameters of a baseline are drawn on the host, from a private `np.random.Philox` stream keyed by
(seed, baseline index) -- the global `np.random` stream the Preprocessor's shuffle uses is left
alone.  Parity with the reference is therefore DISTRIBUTIONAL (tests/test_gpu_synth.py,
tests/test_host_logic.py), not bit-wise: its host MT19937 stream cannot be reproduced on a device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native
from ..utils.device import current_stream_ptr, require_cuda

#: synthetic_val_1k.yaml:15-20
DEFAULT_RFI_COUNTS = {
    "narrowband_persistent": 20, "broadband_persistent": 5, "frequency_sweep": 1,
    "narrowband_intermittent": 0, "narrowband_bursty": 20, "broadband_bursty": 5,
}
MAX_BANDS = 64


def _count(c, rng):
    if isinstance(c, (list, tuple)) and len(c) == 2:      # :566-567
        return int(rng.integers(c[0], c[1] + 1))
    return int(c)


def draw_rfi_events(rng, nc, nt, rfi_counts, rfi_power_min, rfi_power_max):
    """The RFI events of ONE baseline, drawn like `_generate_single_sample` + `_add_*` do
    (same distributions, `rng.integers(lo, hi)` for `np.random.randint(lo, hi)`), returned in
    separable form: (row_amp [nc], col_amp [nt], band_rows [k, 2], band_amp [k, nt],
    sweeps [s, 6], params list)."""
    row_amp = np.zeros(nc, np.float32)
    col_amp = np.zeros(nt, np.float32)
    band_rows, band_amp, sweeps, params = [], [], [], []
    order = ["narrowband_persistent", "broadband_persistent", "narrowband_intermittent",
             "narrowband_bursty", "broadband_bursty", "frequency_sweep"]
    for kind in order:
        for _ in range(_count(rfi_counts.get(kind, 0), rng)):
            amp = float(rng.uniform(rfi_power_min, rfi_power_max) * 1000.0)   # Jy -> mJy, :573
            if kind == "narrowband_persistent":                                 # :675-694
                cf, bw = int(rng.integers(int(nc * 0.1), int(nc * 0.9))), int(rng.integers(1, 10))
                row_amp[max(0, cf - bw // 2):min(nc, cf + bw // 2 + 1)] += amp
                p = {"center_freq": cf, "bandwidth": bw}
            elif kind == "broadband_persistent":                                # :696-709
                ct, tw = int(rng.integers(int(nt * 0.1), int(nt * 0.9))), int(rng.integers(5, 50))
                col_amp[max(0, ct - tw // 2):min(nt, ct + tw // 2)] += amp
                p = {"center_time": ct, "time_width": tw}
            elif kind == "narrowband_intermittent":                             # :711-739
                cf, bw = int(rng.integers(int(nc * 0.1), int(nc * 0.9))), int(rng.integers(2, 15))
                period, duty = int(rng.integers(20, 200)), float(rng.uniform(0.1, 0.5))
                prof = np.zeros(nt, np.float32)
                dur = int(period * duty)
                for t in range(0, nt, period):
                    prof[t:min(nt, t + dur)] = amp
                band_rows.append((max(0, cf - bw // 2), min(nc, cf + bw // 2)))
                band_amp.append(prof)
                p = {"center_freq": cf, "bandwidth": bw, "period": period, "duty_cycle": duty}
            elif kind == "narrowband_bursty":                                   # :741-768
                cf, bw = int(rng.integers(int(nc * 0.1), int(nc * 0.9))), int(rng.integers(2, 20))
                nb = int(rng.integers(3, 15))
                times_ = rng.choice(nt, nb, replace=False)
                widths = rng.integers(2, 20, nb)
                prof = np.zeros(nt, np.float32)
                for t, w in zip(times_, widths):
                    prof[max(0, int(t) - int(w) // 2):min(nt, int(t) + int(w) // 2)] = amp
                band_rows.append((max(0, cf - bw // 2), min(nc, cf + bw // 2)))
                band_amp.append(prof)
                p = {"center_freq": cf, "bandwidth": bw, "num_bursts": nb}
            elif kind == "broadband_bursty":                                    # :770-786
                nb = int(rng.integers(2, 10))
                times_ = rng.choice(nt, nb, replace=False)
                widths = rng.integers(1, 5, nb)
                prof = np.zeros(nt, np.float32)
                for t, w in zip(times_, widths):
                    prof[max(0, int(t) - int(w) // 2):min(nt, int(t) + int(w) // 2)] = amp
                col_amp += prof
                p = {"num_bursts": nb}
            elif kind == "frequency_sweep":                                     # :788-821
                f0 = int(rng.integers(int(nc * 0.1), int(nc * 0.5)))
                f1 = int(rng.integers(int(nc * 0.5), int(nc * 0.9)))
                bw, so = int(rng.integers(2, 10)), int(rng.choice([1, 2]))
                sweeps.append((f0, f1, bw, so, amp, 0.0))
                p = {"start_freq": f0, "end_freq": f1, "bandwidth": bw, "sweep_order": so}
            else:
                continue
            params.append({**p, "type": kind, "amplitude_mjy": amp})
    if len(band_rows) > MAX_BANDS:
        raise ValueError(f"at most {MAX_BANDS} narrow-band bursty / intermittent events per baseline")
    return (row_amp, col_amp, np.asarray(band_rows, np.int32).reshape(-1, 2),
            np.asarray(band_amp, np.float32).reshape(-1, nt), np.asarray(sweeps, np.float32).reshape(-1, 6), params)


class SyntheticDataGenerator:
    """`generate_cube` is `_generate_single_sample` (:520-656) for many baselines at once, on the
    device.  `seed` and `first_baseline` make a cube reproducible and shardable: baseline b of
    the VLA-scale cube is the same array whichever rank generates it."""

    def __init__(self, config=None, device=None):
        self.config = config
        self._device = device

    def generate_cube(self, n_baselines, num_channels, num_times, *, noise_level=1.0, rfi_power_min=1000.0,
                      rfi_power_max=10000.0, rfi_counts=None, enable_bandpass=True, bandpass_order=8,
                      num_polarizations=4, pol_corr=0.8, seed=1234, first_baseline=0, rfi=True):
        """-> (complex64 cube (B, Npol, C, T), bool mask (B, Npol, C, T), rfi_params per baseline)."""
        lib = _native.load()
        device = require_cuda(self._device)
        counts = dict(DEFAULT_RFI_COUNTS if rfi_counts is None else rfi_counts) if rfi else {}
        nc, nt, nb = int(num_channels), int(num_times), int(n_baselines)
        ev = []
        for b in range(nb):
            rng = np.random.Generator(np.random.Philox(np.random.SeedSequence([int(seed), int(first_baseline) + b])))
            ev.append(draw_rfi_events(rng, nc, nt, counts, rfi_power_min, rfi_power_max))
        n_bands = max([e[2].shape[0] for e in ev], default=0)
        n_sweeps = max([e[4].shape[0] for e in ev], default=0)
        row = np.stack([e[0] for e in ev]) if nb else np.zeros((0, nc), np.float32)
        col = np.stack([e[1] for e in ev]) if nb else np.zeros((0, nt), np.float32)
        rows = np.zeros((nb, max(n_bands, 1), 2), np.int32)      # unused bands: empty row range
        amps = np.zeros((nb, max(n_bands, 1), nt), np.float32)
        sw = np.zeros((nb, max(n_sweeps, 1), 6), np.float32)      # unused sweeps: amplitude 0, width 0
        for b, e in enumerate(ev):
            rows[b, :e[2].shape[0]] = e[2]
            amps[b, :e[3].shape[0]] = e[3]
            sw[b, :e[4].shape[0]] = e[4]
        sp = _native.RfiSynth(channels=nc, times=nt, n_pol=int(num_polarizations), enable_bandpass=int(bool(enable_bandpass)),
                              bandpass_order=int(bandpass_order), n_bands=n_bands, n_sweeps=n_sweeps,
                              noise_level=float(noise_level), pol_corr=float(pol_corr), seed=int(seed) & (2**64 - 1))
        with torch.cuda.device(device):
            d = [torch.from_numpy(a).to(device) for a in (row, col, rows, amps, sw)]
            cube = torch.empty((nb, int(num_polarizations), nc, nt), dtype=torch.complex64, device=device)
            mask = torch.empty(cube.shape, dtype=torch.uint8, device=device)
            # at most 65535 baselines per launch (grid.z)
            rc = lib.rfi_synth_waterfalls(C.byref(sp), nb, int(first_baseline), d[0].data_ptr(), d[1].data_ptr(),
                                          d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                                          cube.data_ptr(), mask.data_ptr(), current_stream_ptr(device))
            _native.check(rc, "rfi_synth_waterfalls")
        return cube, mask.view(torch.bool), [e[5] for e in ev]
