"""Full-size runs of the BASELINE.json configurations on one GPU, checked through
size-independent properties (the oracle cannot finish these sizes): kept-patch bookkeeping against
the per-tile statistics, a checksum of checksums (labels summed per patch == the tile's flag count,
for every rotation), rotation consistency of the written patches, min-max invariants of the image
channels, determinism, metric identities, shard additivity -- plus oracle parity on a one-baseline
slice at the full waterfall size ("parity ... on a slice copied to the host").
These sizes also cross 2^32 output elements, i.e. they exercise the 64-bit index arithmetic."""
import numpy as np
import pytest
import torch

import oracle
from tests.test_gpu_bigtile import _routes
from tests.test_gpu_parity import IMG_ATOL, IMG_RTOL, image_errors

pytestmark = pytest.mark.gpu

NB1, NB2 = (0.0 - 0.456) / 0.224, (0.0 - 0.406) / 0.225      # ImageNet-normalised 0 of channels 1, 2
ONE1 = (1.0 - 0.456) / 0.224


def _cube(n_bl, channels, times, seed=1234):
    from rfi_toolbox_b200 import SyntheticDataGenerator
    cube, mask, _ = SyntheticDataGenerator().generate_cube(n_bl, channels, times, seed=seed)
    return cube, mask


def _check_properties(pre, ds, P, R, nh, nw, n_wf, sample=64):
    st = _routes(pre)
    n_tiles = n_wf * nh * nw
    assert len(st) == n_tiles
    keep_tile = st["n_flagged"] > 0
    order = np.asarray(pre.order)
    # ---- bookkeeping: exactly the R rotations of every tile with a flag, each once
    assert len(ds) == R * int(keep_tile.sum()) == len(order)
    assert len(np.unique(order)) == len(order)
    per = nh * nw
    w, rem = np.divmod(order, R * per)
    r, t = np.divmod(rem, per)
    ti = np.where(r <= 1, t // nw, t % nh)
    tj = np.where(r <= 1, t % nw, t // nh)
    ti = np.where(r == 1, nh - 1 - ti, ti)
    tj = np.where(r == 3, nw - 1 - tj, tj)
    tile = w * per + ti * nw + tj                              # original tile of every output patch
    assert keep_tile[tile].all()
    assert np.array_equal(np.bincount(tile, minlength=n_tiles), R * keep_tile.astype(np.int64))
    # ---- checksum of checksums: labels of patch k sum to the flag count of its tile
    sums = ds.labels.view(len(ds), -1).sum(dim=1, dtype=torch.int64).cpu().numpy()
    assert np.array_equal(sums, st["n_flagged"][tile])
    assert int(sums.sum()) == int((st["n_flagged"][keep_tile].astype(np.int64) * R).sum())
    # ---- rotation consistency on a sample of tiles: r1 = flipud(r0), r2 = r0.T, r3 = flipud(r0.T)
    slot = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(tile, r))}
    rng = np.random.default_rng(0)
    for tl in rng.choice(np.flatnonzero(keep_tile), size=min(sample, int(keep_tile.sum())), replace=False):
        l0 = ds.labels[slot[(int(tl), 0)]]
        i0 = ds.images[slot[(int(tl), 0)]]
        if R >= 2:
            assert torch.equal(ds.labels[slot[(int(tl), 1)]], torch.flip(l0, dims=[0]))
            assert torch.allclose(ds.images[slot[(int(tl), 1)]][..., 1], torch.flip(i0[..., 1], dims=[0]), rtol=0, atol=1e-6)
        if R >= 4:
            assert torch.equal(ds.labels[slot[(int(tl), 2)]], l0.T)
            assert torch.equal(ds.labels[slot[(int(tl), 3)]], torch.flip(l0.T, dims=[0]))
            assert torch.allclose(ds.images[slot[(int(tl), 2)]][..., 1], i0[..., 1].T, rtol=0, atol=1e-6)
    # ---- image invariants: channel 2 constant; channel 1 min-max normalised per patch; all finite
    assert bool((ds.images[..., 2] == NB2).all())
    c1 = ds.images[..., 1].reshape(len(ds), -1)
    lo, hi = c1.min(dim=1).values, c1.max(dim=1).values
    flat = hi == lo                                           # constant patches map to zeros (:157-163)
    assert torch.allclose(lo[~flat], torch.full_like(lo[~flat], NB1), atol=2e-5)
    assert torch.allclose(hi[~flat], torch.full_like(hi[~flat], ONE1), atol=2e-5)
    assert bool(torch.isfinite(ds.images[..., 0]).all())
    return st, tile


def _slice_parity(cube, b, kw, P, max_label_mismatch=0.0, label=""):
    """oracle vs GPU on one baseline at the full waterfall size.  `max_label_mismatch` > 0: the
    stretch is LOG10 (host-dependent upstream of the thresholds), mismatches are counted."""
    from rfi_toolbox_b200 import Preprocessor
    sl = cube[b:b + 1]
    np.random.seed(3)
    pre = Preprocessor(sl, None, magnitude=True)
    ds = pre.create_dataset(**kw)
    np.random.seed(3)
    ods, inter = oracle.create_dataset(np.abs(sl.cpu().numpy()), None, return_intermediates=True, **kw)
    assert np.array_equal(pre.order, inter["order"])
    labs, imgs = ds.labels.cpu().numpy(), ds.images.cpu().numpy()
    diff = np.abs(imgs - ods.images)
    print(f"[slice parity] {label}: {len(ods.images)} patches, label mismatch fraction {(labs != ods.labels).mean():.3e}, "
          f"max |cuda - numpy| {np.nanmax(diff):.3e}")
    if max_label_mismatch == 0.0:
        assert np.array_equal(labs, ods.labels)
        assert np.allclose(imgs, ods.images, rtol=IMG_RTOL, atol=IMG_ATOL, equal_nan=True)
        image_errors(imgs, ods, inter, label, max_patches=256)
    else:
        assert (labs != ods.labels).mean() <= max_label_mismatch
        ok = np.isclose(imgs, ods.images, rtol=IMG_RTOL, atol=IMG_ATOL, equal_nan=True)
        assert (~ok).mean() <= 10 * max_label_mismatch + 1e-4


def test_config2_full_size_properties(native_lib):
    """configs[1]: 45 baselines x 4 pols x 1024 x 1024, SQRT, MAD sigma 5, 4 rotations."""
    from rfi_toolbox_b200 import Preprocessor, evaluate_segmentation
    from rfi_toolbox_b200.evaluation.metrics import confusion_counts
    cube, mask = _cube(45, 1024, 1024)
    kw = dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    np.random.seed(0)
    pre = Preprocessor(cube, None, magnitude=True)
    ds = pre.create_dataset(**kw)
    _check_properties(pre, ds, 128, 4, 8, 8, 45 * 4)
    # determinism: same seed, same bits
    np.random.seed(0)
    pre2 = Preprocessor(cube, None, magnitude=True)
    ds2 = pre2.create_dataset(**kw)
    assert np.array_equal(pre.order, pre2.order) and torch.equal(ds.labels, ds2.labels) and torch.equal(ds.images, ds2.images)
    del ds2, pre2
    # metric identities and shard additivity
    m = evaluate_segmentation(ds.labels, ds.labels)
    assert all(v == 1.0 for v in m.values())
    truth = ds.labels ^ (torch.rand(ds.labels.shape, device=ds.labels.device) < 0.01).to(torch.uint8)
    tp, fp, fn = confusion_counts(ds.labels, truth)
    assert tp + fp == int(ds.labels.sum()) and tp + fn == int(truth.sum())
    half = len(ds) // 2
    a, b = confusion_counts(ds.labels[:half], truth[:half]), confusion_counts(ds.labels[half:], truth[half:])
    assert (tp, fp, fn) == tuple(x + y for x, y in zip(a, b))
    _slice_parity(cube, 7, kw, 128, label="configs[1] baseline 7")


def test_config3_half_shard_properties(native_lib):
    """configs[2]: half of one GPU's shard of the VLA-scale cube (22 baselines x 4 x 4096 x 2048),
    LOG10 stretch -- 8.8e9 output floats; the exact-zero bandpass rows take the general route."""
    from rfi_toolbox_b200 import Preprocessor
    cube, _ = _cube(22, 4096, 2048, seed=77)
    kw = dict(patch_size=128, stretch="LOG10", flag_sigma=5, use_custom_flags=False)
    np.random.seed(0)
    pre = Preprocessor(cube, None, magnitude=True)
    ds = pre.create_dataset(**kw)
    st, _ = _check_properties(pre, ds, 128, 4, 32, 16, 22 * 4)
    assert ds.images.numel() > 2**32
    route = st["route"].reshape(22 * 4, 32, 16)
    # pol 0 carries the exact-zero bandpass edge rows (log10(0) = -inf -> MAD fill): general algorithm there,
    # which also finds the raw thresholds for phase 2's fast route (GENERAL | RAW_THRESHOLDS | RAW_FILL = 11);
    # raw thresholds in the interior
    assert ((route[0::4, 0] & 11) == 11).all() and ((route[0::4, 31] & 11) == 11).all()
    assert (route[:, 1:31] & 1).mean() > 0.95
    del ds, pre
    # the headline configuration's literal call against the oracle: one baseline at 4096 x 2048
    _slice_parity(cube, 5, kw, 128, max_label_mismatch=1e-4, label="configs[2] baseline 5")


def test_config5_chunk_properties(native_lib):
    """configs[4]: a 4-baseline chunk of the long-track cube (4 x 4 x 1024 x 16384), P = 256,
    MAD sigma 3, no stretch -- the big-tile path; parity on a slice of two polarisations."""
    from rfi_toolbox_b200 import Preprocessor
    cube, _ = _cube(4, 1024, 16384, seed=5)
    kw = dict(patch_size=256, stretch=None, flag_sigma=3, use_custom_flags=False)
    np.random.seed(0)
    pre = Preprocessor(cube, None, magnitude=True)
    ds = pre.create_dataset(**kw)
    st, _ = _check_properties(pre, ds, 256, 4, 4, 64, 4 * 4, sample=24)
    assert (st["route"] & 1).mean() > 0.99
    del ds, pre
    _slice_parity(cube[:, :2], 1, kw, 256, label="configs[4] baseline 1, 2 pols")


def test_config4_pair_sweep_properties(native_lib):
    """configs[3]: IoU / F1 / FFI sweep over 100 000 predicted / ground-truth 128 x 128 pairs."""
    from rfi_toolbox_b200 import compute_ffi, compute_ffi_batch, evaluate_segmentation, evaluate_segmentation_batch
    n = 100_000
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    true = torch.rand((n, 128, 128), generator=g, device=dev) < 0.10
    pred = true ^ (torch.rand((n, 128, 128), generator=g, device=dev) < 0.02)
    data = torch.view_as_complex(torch.randn((n, 128, 128, 2), generator=g, device=dev))
    data = data * (1.0 + 99.0 * true)
    m = evaluate_segmentation_batch(pred, true)
    tot = evaluate_segmentation(pred, true)
    tp, fp, fn = int(m["tp"].sum()), int(m["fp"].sum()), int(m["fn"].sum())
    assert np.isclose(tot["iou"], tp / (tp + fp + fn)) and np.isclose(tot["dice"], 2 * tp / (2 * tp + fp + fn))
    f = compute_ffi_batch(data, pred)
    assert len(f["ffi"]) == n and np.isfinite(f["ffi"]).all()
    for i in np.random.default_rng(0).choice(n, 12, replace=False):     # batch == per-pair calls == oracle
        one = evaluate_segmentation(pred[i], true[i])
        assert all(np.isclose(one[k], m[k][i], rtol=0, atol=0) for k in one)
        fi = compute_ffi(data[i], pred[i])
        want = oracle.compute_ffi(data[i].cpu().numpy(), pred[i].cpu().numpy())
        for k in want:
            assert abs(f[k][i] - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, f[k][i], want[k])
            assert abs(fi[k] - want[k]) <= 1e-6 * max(1.0, abs(want[k]))
