"""BASELINE config 4 as one launch (`rfi_pair_sweep` / `evaluate_pairs`): per-pair compute_ffi +
evaluate_segmentation against the oracle called pair by pair.

Bars: TP / FP / FN and hence the five ratios bit-exact (integer counts, the reference's float64
formulas); medians and MADs bit-exact (exact order statistics, whichever of the sampled-bracket or
radix routes a pair takes); means / stds / FFI within 1e-6 relative (NumPy sums pairwise in float32)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _pairs(n, dtype, seed=0, shape=(128, 128), p_true=0.10):
    rng = np.random.default_rng(seed)
    truth = rng.random((n,) + shape) < p_true
    data = (rng.normal(0, 1, (n,) + shape) + 1j * rng.normal(0, 1, (n,) + shape)) * np.where(truth, 100.0, 1.0)
    if np.dtype(dtype).kind != "c":
        data = np.abs(data)
    pred = truth ^ (rng.random(truth.shape) < 0.02)
    return data.astype(dtype), truth, pred


def _check(data, pred, truth, got, stats_before=None, stats_after=None, indices=None):
    for i in (range(len(data)) if indices is None else indices):
        try:
            want = oracle.compute_ffi(data[i], pred[i])
        except ZeroDivisionError:
            assert np.isnan(got["ffi"][i]), i
            want = None
        if want is not None:
            for k, v in want.items():
                assert got[k][i] == pytest.approx(v, rel=1e-6, abs=1e-9, nan_ok=True), (i, k, got[k][i], v)
        m = oracle.evaluate_segmentation(pred[i], truth[i])
        for k, v in m.items():
            assert got[k][i] == v, (i, k, got[k][i], v)
        for st, fl in ((stats_before, None), (stats_after, pred[i])):
            if st is None:
                continue
            ws = oracle.compute_statistics(data[i], fl)
            for k in ("median", "mad"):
                a, b = st[k][i], ws[k]
                assert (np.isnan(a) and np.isnan(b)) or np.float32(a) == np.float32(b), (i, k, a, b)
            # NumPy sums pairwise in float32: its own error is ~1e-7 of mean(|x|), which for signed data
            # (mean near zero) is far more than 1e-6 of the mean itself
            kept = np.abs(data[i]) if fl is None else np.abs(data[i][~fl])
            scale = float(np.nanmean(kept)) if kept.size and np.isfinite(kept).any() else 0.0
            scale = scale if np.isfinite(scale) else 0.0
            for k in ("mean", "std"):
                assert st[k][i] == pytest.approx(ws[k], rel=1e-6, abs=1e-6 * scale, nan_ok=True), (i, k)
            assert int(st["count"][i]) == ws["count"]


@pytest.mark.parametrize("dtype", [np.complex64, np.float32])
def test_pair_sweep_matches_per_pair_oracle(native_lib, dtype):
    from rfi_toolbox_b200 import compute_statistics_batch, evaluate_pairs
    n = 40
    data, truth, pred = _pairs(n, dtype, seed=2)
    pred[3] = True                      # all flagged -> the reference's guard branch
    data[5, 7, 9] = np.nan              # NaN in the data ...
    pred[5, 7, 9] = False               # ... left unflagged -> guard branch too
    data[6, 1, 1] = np.nan
    pred[6, 1, 1] = True                # NaN flagged away: only `before` is NaN
    pred[7] = False                     # nothing flagged: after == before, reductions 0
    truth[8] = False                    # no true pixel
    pred[9] = False; truth[9] = False   # empty pair: iou 1, precision 1, recall 1
    data[10, 3, 4] = np.inf             # one infinite sample (radix route)
    data[11] = np.round(np.abs(data[11]) * 2) / 2      # heavy duplicates: brackets of equal keys
    data[12, :64] = 7.0                 # half of the pair one constant
    if dtype == np.float32:
        data[13] *= np.where(np.random.default_rng(1).random(data[13].shape) < 0.5, -1.0, 1.0)   # signed real data
        data[14, 0, 0] = -np.inf
    got = evaluate_pairs(data, pred, truth, errors="nan")
    sb = compute_statistics_batch(data)
    sa = compute_statistics_batch(data, pred)
    _check(data, pred, truth, got, sb, sa)


@pytest.mark.parametrize("shape", [(50, 37), (33, 31), (3, 5), (128, 127), (1, 1)])
def test_pair_sweep_ragged_and_small_pairs(native_lib, shape):
    """Pair sizes that are not multiples of 4 (scalar loads, unaligned pair starts), smaller than the
    sample (radix route) and down to one sample."""
    from rfi_toolbox_b200 import compute_statistics_batch, evaluate_pairs
    data, truth, pred = _pairs(7, np.complex64, seed=4, shape=shape, p_true=0.3)
    got = evaluate_pairs(data, pred, truth, errors="nan")
    _check(data, pred, truth, got, compute_statistics_batch(data), compute_statistics_batch(data, pred))


def test_pair_sweep_constant_data_raises(native_lib):
    from rfi_toolbox_b200 import evaluate_pairs
    const = np.ones((2, 16, 16), dtype=np.float32)
    z = np.zeros_like(const, dtype=bool)
    with pytest.raises(ZeroDivisionError):
        evaluate_pairs(const, z, z)
    got = evaluate_pairs(const, z, z, errors="nan")
    assert np.isnan(got["ffi"]).all() and (got["iou"] == 1.0).all()
    with pytest.raises(IndexError):
        evaluate_pairs(const, np.zeros_like(const, dtype=np.uint8), z)


def test_pair_sweep_is_deterministic_and_matches_separate_calls(native_lib):
    import torch
    from rfi_toolbox_b200 import compute_ffi_batch, evaluate_pairs, evaluate_segmentation_batch
    n = 3000
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(7)
    true = torch.rand((n, 128, 128), generator=g, device=dev) < 0.10
    pred = true ^ (torch.rand((n, 128, 128), generator=g, device=dev) < 0.02)
    data = torch.view_as_complex(torch.randn((n, 128, 128, 2), generator=g, device=dev))
    data = (data * (1.0 + 99.0 * true)).contiguous()
    a = evaluate_pairs(data, pred, true)
    b = evaluate_pairs(data, pred, true)
    f = compute_ffi_batch(data, pred)
    m = evaluate_segmentation_batch(pred, true)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    for k in f:
        assert np.array_equal(a[k], f[k]), k
    for k in ("iou", "precision", "recall", "f1", "dice", "tp", "fp", "fn"):
        assert np.array_equal(a[k], m[k]), k
    idx = np.random.default_rng(0).choice(n, 10, replace=False)
    _check(data.cpu().numpy(), pred.cpu().numpy(), true.cpu().numpy(), a, indices=idx)
