"""Synthetic generator (rfi_toolbox_b200.data_generation; the step before the hot path).
Parity with the reference generator is DISTRIBUTIONAL: its host MT19937 stream cannot be reproduced
on a device.  CPU tests: the host-side event draws (ranges, counts, separable form) and, with the
live reference, the flagged-pixel fraction.  GPU tests: the pixel kernel's properties."""
import numpy as np
import pytest

from tests.conftest import REFERENCE

from rfi_toolbox_b200.data_generation import DEFAULT_RFI_COUNTS, draw_rfi_events


def _rasterise(ev, nc, nt):
    """signal / mask of one baseline from the separable event form (CPU restatement of the kernel)."""
    row, col, rows, amps, sweeps, _ = ev
    sig = row[:, None].astype(np.float64) + col[None, :]
    for (r0, r1), prof in zip(rows, amps):
        sig[r0:r1, :] += prof[None, :]
    t = np.arange(nt, dtype=np.float32)
    for f0, f1, bw, so, amp, _ in sweeps:
        prog = t / np.float32(nt)
        if so == 2:
            prog = prog * prog
        centre = (np.float32(f0) + (np.float32(f1) - np.float32(f0)) * prog).astype(np.int64)
        for tt in range(nt):
            lo, hi = max(0, centre[tt] - int(bw) // 2), min(nc, centre[tt] + int(bw) // 2)
            sig[lo:hi, tt] += amp
    return sig, sig > 0


def _events(seed, b, nc, nt):
    rng = np.random.Generator(np.random.Philox(np.random.SeedSequence([seed, b])))
    return draw_rfi_events(rng, nc, nt, DEFAULT_RFI_COUNTS, 1000.0, 10000.0)


def test_event_draws_follow_the_reference_ranges():
    nc, nt = 512, 768
    ev = _events(7, 0, nc, nt)
    params = ev[5]
    assert len(params) == sum(DEFAULT_RFI_COUNTS.values())                 # synthetic_val_1k.yaml:15-20
    assert all(1e6 <= p["amplitude_mjy"] <= 1e7 for p in params)           # U(1000, 10000) Jy in mJy
    for p in params:
        if p["type"] == "narrowband_persistent":
            assert int(nc * 0.1) <= p["center_freq"] < int(nc * 0.9) and 1 <= p["bandwidth"] < 10
        if p["type"] == "broadband_persistent":
            assert int(nt * 0.1) <= p["center_time"] < int(nt * 0.9) and 5 <= p["time_width"] < 50
        if p["type"] == "narrowband_bursty":
            assert 2 <= p["bandwidth"] < 20 and 3 <= p["num_bursts"] < 15
        if p["type"] == "frequency_sweep":
            assert p["start_freq"] < int(nc * 0.5) <= p["end_freq"] and p["sweep_order"] in (1, 2)
    assert ev[2].shape == (20, 2) and ev[3].shape == (20, nt) and ev[4].shape == (1, 6)
    # deterministic in (seed, baseline), different across baselines
    again = _events(7, 0, nc, nt)
    assert all(np.array_equal(a, b) for a, b in zip(ev[:5], again[:5]))
    assert not np.array_equal(ev[0], _events(7, 1, nc, nt)[0])


@pytest.mark.reference
def test_flagged_fraction_matches_the_reference_generator():
    """Same recipe -> same coverage statistics as SyntheticDataGenerator._generate_single_sample."""
    import sys
    sys.path.insert(0, str(REFERENCE))
    from rfi_toolbox.data_generation.synthetic_generator import SyntheticDataGenerator as Ref
    nc = nt = 256
    ref = Ref.__new__(Ref)
    cfg = {k: {"count": v} for k, v in DEFAULT_RFI_COUNTS.items()}
    np.random.seed(3)
    fr_ref, fr_ours, amp_ref, amp_ours = [], [], [], []
    for b in range(24):
        _, m, params = ref._generate_single_sample(nc, nt, 1.0, 1000.0, 10000.0, cfg, True, 8, 4, 0.8, {})
        fr_ref.append(m[0, 0].mean())
        amp_ref += [p["amplitude_mjy"] for p in params]
        ev = _events(11, b, nc, nt)
        fr_ours.append(_rasterise(ev, nc, nt)[1].mean())
        amp_ours += [p["amplitude_mjy"] for p in ev[5]]
    fr_ref, fr_ours = np.array(fr_ref), np.array(fr_ours)
    se = np.sqrt(fr_ref.var() / len(fr_ref) + fr_ours.var() / len(fr_ours))
    assert abs(fr_ref.mean() - fr_ours.mean()) < 4 * se + 0.01, (fr_ref.mean(), fr_ours.mean())
    assert abs(np.mean(amp_ref) - np.mean(amp_ours)) < 0.05 * np.mean(amp_ref)


@pytest.mark.gpu
def test_cube_properties(native_lib):
    import torch
    from rfi_toolbox_b200.data_generation import SyntheticDataGenerator
    gen = SyntheticDataGenerator()
    nb, nc, nt = 3, 512, 768
    cube, mask, params = gen.generate_cube(nb, nc, nt, seed=1234)
    torch.cuda.synchronize()
    assert cube.shape == (nb, 4, nc, nt) and cube.dtype == torch.complex64
    assert mask.shape == cube.shape and mask.dtype == torch.bool and len(params) == nb
    z, m = cube.cpu().numpy(), mask.cpu().numpy()
    # mask and RFI signal = the CPU rasterisation of the same events
    for b in range(nb):
        sig, mk = _rasterise(_events(1234, b, nc, nt), nc, nt)
        assert np.array_equal(m[b, 0], mk) and np.array_equal(m[b, 1], mk)
        assert not m[b, 2].any() and not m[b, 3].any()
        a0, a1 = np.abs(z[b, 0]), np.abs(z[b, 1])
        hot = mk & (sig > 1e5)
        assert np.allclose(a0[hot], sig[hot], rtol=1e-4, atol=2.0)               # base ~ 1 on 1e6..1e7
        assert np.allclose(a1[hot], 0.8 * sig[hot], rtol=1e-4, atol=2.0)         # pol_corr (:622-627)
    # clean noise: N(1, 0.1) in the flat part of the band; bandpass edge rows exactly zero (:657-673)
    mid = slice(int(nc * 0.1), int(nc * 0.9))
    clean = np.abs(z[:, 0, mid, :])[~m[:, 0, mid, :]]
    assert abs(clean.mean() - 1.0) < 2e-3 and abs(clean.std() - 0.1) < 2e-3
    # (pol 1 adds (1 - corr) * N(0, 0.1) on top of the band-passed base, :622-627: not zero there)
    assert (z[:, 0, 0, :][~m[:, 0, 0, :]] == 0).all() and (z[:, 0, -1, :][~m[:, 0, -1, :]] == 0).all()
    n23 = np.abs(z[:, 2:])                                                        # noise only, no bandpass
    assert abs(n23.mean() - 1.0) < 1e-3 and abs(n23.std() - 0.1) < 1e-3
    ph = np.angle(z[:, 2])
    assert abs(np.cos(ph).mean()) < 3e-3 and abs(np.sin(ph).mean()) < 3e-3        # uniform phase (:639)
    # pols are not copies of each other
    assert abs(np.corrcoef(n23[:, 0].ravel()[:100000], n23[:, 1].ravel()[:100000])[0, 1]) < 0.02


@pytest.mark.gpu
def test_cube_is_a_pure_function_of_seed_and_baseline(native_lib):
    """Sharding invariance: baseline b is the same array whichever rank / call generates it."""
    import torch
    from rfi_toolbox_b200.data_generation import SyntheticDataGenerator
    gen = SyntheticDataGenerator()
    full, mfull, _ = gen.generate_cube(4, 256, 384, seed=99)
    part, mpart, _ = gen.generate_cube(2, 256, 384, seed=99, first_baseline=2)
    other, _, _ = gen.generate_cube(2, 256, 384, seed=100, first_baseline=2)
    torch.cuda.synchronize()
    assert torch.equal(full[2:], part) and torch.equal(mfull[2:], mpart)
    assert not torch.equal(part, other)
    # 2 polarisations, no RFI, no bandpass
    c2, m2, p2 = gen.generate_cube(1, 128, 128, num_polarizations=2, rfi=False, enable_bandpass=False)
    assert c2.shape == (1, 2, 128, 128) and not m2.any() and p2 == [[]]
    assert abs(float(c2.abs().mean()) - 1.0) < 5e-3


@pytest.mark.gpu
def test_generator_feeds_the_preprocessor(native_lib):
    """The generator's own call of the path (synthetic_generator.py:92-105): custom flags."""
    import torch
    from rfi_toolbox_b200 import Preprocessor
    from rfi_toolbox_b200.data_generation import SyntheticDataGenerator
    cube, mask, _ = SyntheticDataGenerator().generate_cube(2, 256, 256, seed=5)
    np.random.seed(0)
    ds = Preprocessor(cube, mask).create_dataset(patch_size=128, stretch=None, use_custom_flags=True,
                                                 normalize_before_stretch=False, num_workers=0)
    torch.cuda.synchronize()
    assert ds.images.shape[1:] == (128, 128, 3) and len(ds) > 0
    assert int(ds.labels.sum()) > 0 and torch.isfinite(ds.images).all()
