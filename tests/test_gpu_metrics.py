"""GPU parity for evaluate_segmentation / compute_statistics / compute_ffi."""
import numpy as np
import pytest
import torch

import oracle
from tests.cubes import make_cube

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1,), (5,), (64, 64), (3, 128, 128), (1000003,)])
@pytest.mark.parametrize("pd,td", [(np.bool_, np.bool_), (np.uint8, np.bool_), (np.float32, np.uint8),
                                   (np.float64, np.float32), (np.int64, np.int16), (np.float16, np.bool_)])
def test_confusion_counts_bit_exact(native_lib, shape, pd, td):
    from rfi_toolbox_b200 import evaluate_segmentation
    from rfi_toolbox_b200.evaluation import confusion_counts
    rng = np.random.default_rng(1)
    p = (rng.random(shape) < 0.3).astype(pd)
    t = (rng.random(shape) < 0.25).astype(td)
    if np.dtype(pd).kind == "f":
        flat = p.reshape(-1)
        flat[::7] = np.nan   # NaN is truthy
        flat[1::11] = -0.0   # signed zero is falsy
    assert confusion_counts(p, t) == oracle.confusion_counts(p, t)
    a, b = evaluate_segmentation(p, t), oracle.evaluate_segmentation(p, t)
    assert a == b and all(type(a[k]) is type(b[k]) for k in a)


def test_metrics_guards_and_torch_inputs(native_lib):
    from rfi_toolbox_b200 import evaluate_segmentation
    z = np.zeros((32, 32), dtype=bool)
    o = np.ones((32, 32), dtype=bool)
    for p, t in ((z, z), (z, o), (o, z), (o, o)):
        assert evaluate_segmentation(p, t) == oracle.evaluate_segmentation(p, t)
    p = torch.rand(4, 128, 128, device="cuda") > 0.5
    t = torch.rand(4, 128, 128, device="cuda") > 0.5
    assert evaluate_segmentation(p, t) == oracle.evaluate_segmentation(p.cpu().numpy(), t.cpu().numpy())
    # misaligned views
    pn, tn = p.cpu().numpy().reshape(-1)[3:-1], t.cpu().numpy().reshape(-1)[3:-1]
    assert evaluate_segmentation(torch.from_numpy(pn.copy()).cuda(), tn) == oracle.evaluate_segmentation(pn, tn)


def test_segmented_counts(native_lib):
    from rfi_toolbox_b200.evaluation import evaluate_segmentation_batch
    rng = np.random.default_rng(2)
    t = rng.random((257, 128, 128)) < 0.1
    p = t ^ (rng.random(t.shape) < 0.02)
    p[3] = False; t[3] = False
    p[4] = False
    got = evaluate_segmentation_batch(p, t)
    for i in range(len(p)):
        ref = oracle.evaluate_segmentation(p[i], t[i])
        for k in ref:
            assert got[k][i] == ref[k], (i, k)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.float64, np.complex128])
def test_statistics_and_ffi(native_lib, dtype):
    from rfi_toolbox_b200 import compute_ffi, compute_statistics
    data, mask = make_cube(dtype=dtype, seed=31)
    for flags in (None, mask):
        a, b = compute_statistics(data, flags), oracle.compute_statistics(data, flags)
        assert a["count"] == b["count"] and a["flagged_fraction"] == b["flagged_fraction"]
        assert a["median"] == b["median"], "median is an exact order statistic"
        assert a["mad"] == b["mad"], "MAD is an exact order statistic"
        for k in ("mean", "std"):
            assert a[k] == pytest.approx(b[k], rel=1e-6)
    a, b = compute_ffi(data, mask), oracle.compute_ffi(data, mask)
    for k in b:
        assert a[k] == pytest.approx(b[k], rel=1e-6, abs=1e-9)
    assert compute_ffi(data, np.ones_like(mask)) == oracle.compute_ffi(data, np.ones_like(mask))
    with pytest.raises(IndexError):
        compute_ffi(data, mask.astype(np.uint8))
    # a mask over the leading dimensions only (NumPy's boolean-mask rule: whole waterfalls / baselines dropped)
    for lead in (mask[:, :, 0, 0].copy(), mask[:, 0, 0, 0].copy()):
        lead.flat[0] = True
        lead.flat[-1] = False
        a, b = compute_statistics(data, lead), oracle.compute_statistics(data, lead)
        assert a["count"] == b["count"] and a["median"] == b["median"] and a["mad"] == b["mad"]
    with pytest.raises(IndexError):
        compute_statistics(data, mask[:, :, :, :-1])


def test_sqrt_unit_range_is_exact(native_lib):
    """cabs_fast's square root (no range test, argument in [1, 2]) equals sqrt.rn.f32 on every
    one of the 2^23 + 1 possible arguments."""
    import torch
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = native_lib.rfi_selftest_sqrt_unit(bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert int(bad.item()) == 0


def test_cabs_fast_is_exact(native_lib):
    """The magnitude of the hot path (own quotient sequence, ONE range test per sample) equals the IEEE
    chain np.abs runs (div.rn, fma, sqrt.rn, mul) on 2^32 operand pairs: arbitrary bit patterns, close
    exponents, and the edges of the fast range."""
    import torch
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = native_lib.rfi_selftest_cabs_fast(bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert int(bad.item()) == 0


def _pairs(n, dtype, seed=0, shape=(128, 128)):
    rng = np.random.default_rng(seed)
    truth = rng.random((n,) + shape) < 0.10
    data = (rng.normal(0, 1, (n,) + shape) + 1j * rng.normal(0, 1, (n,) + shape)) * np.where(truth, 100.0, 1.0)
    if np.dtype(dtype).kind != "c":
        data = np.abs(data)
    pred = truth ^ (rng.random(truth.shape) < 0.02)
    return data.astype(dtype), truth, pred


@pytest.mark.parametrize("dtype", [np.complex64, np.float32, np.complex128, np.float64])
def test_ffi_sweep_matches_per_pair_oracle(native_lib, dtype):
    """BASELINE config 4 (scaled down): compute_ffi / compute_statistics over a stack of
    128 x 128 pairs in one launch == the reference called pair by pair.  Medians and MADs are
    exact order statistics (bit-exact); means / stds within 1e-6 relative."""
    import oracle
    from rfi_toolbox_b200 import compute_ffi_batch, compute_statistics_batch
    n = 24
    data, truth, pred = _pairs(n, dtype, seed=2)
    pred[3] = True                      # all flagged -> the reference's guard branch
    data[5, 7, 9] = np.nan              # NaN in the data
    pred[5, 7, 9] = False               # ... left unflagged -> guard branch too
    data[6, 1, 1] = np.nan
    pred[6, 1, 1] = True                # NaN flagged away: only `before` is NaN
    got = compute_ffi_batch(data, pred)
    st_b = compute_statistics_batch(data)
    st_a = compute_statistics_batch(data, pred)
    for i in range(n):
        want = oracle.compute_ffi(data[i], pred[i])
        for k, v in want.items():
            assert got[k][i] == pytest.approx(v, rel=1e-6, abs=1e-9, nan_ok=True), (i, k)
        for st, fl in ((st_b, None), (st_a, pred[i])):
            ws = oracle.compute_statistics(data[i], fl)
            for k in ("median", "mad"):
                a, b = st[k][i], ws[k]
                assert (np.isnan(a) and np.isnan(b)) or np.dtype(dtype).type(0).real.dtype.type(a) == b, (i, k, a, b)
            for k in ("mean", "std", "flagged_fraction"):
                assert st[k][i] == pytest.approx(ws[k], rel=1e-6, nan_ok=True), (i, k)
            assert int(st["count"][i]) == ws["count"]


def test_ffi_sweep_ragged_segment_and_errors(native_lib):
    import oracle
    from rfi_toolbox_b200 import compute_ffi_batch
    data, truth, pred = _pairs(5, np.complex64, seed=4, shape=(50, 37))   # segment not a multiple of 512
    got = compute_ffi_batch(data, pred)
    for i in range(5):
        want = oracle.compute_ffi(data[i], pred[i])
        for k, v in want.items():
            assert got[k][i] == pytest.approx(v, rel=1e-6, abs=1e-9)
    const = np.ones((2, 16, 16), dtype=np.float32)
    with pytest.raises(ZeroDivisionError):
        compute_ffi_batch(const, np.zeros_like(const, dtype=bool))
    assert np.isnan(compute_ffi_batch(const, np.zeros_like(const, dtype=bool), errors="nan")["ffi"]).all()
    with pytest.raises(IndexError):
        compute_ffi_batch(const, np.zeros_like(const, dtype=np.uint8))
    # pairs beyond 128 x 128 go through the whole-cube routine, pair by pair
    big, _, bpred = _pairs(2, np.complex64, seed=6, shape=(256, 200))
    got = compute_ffi_batch(big, bpred)
    for i in range(2):
        want = oracle.compute_ffi(big[i], bpred[i])
        for k, v in want.items():
            assert got[k][i] == pytest.approx(v, rel=1e-6, abs=1e-9)


@pytest.mark.parametrize("dtype", [np.complex64, np.float32, np.float64])
def test_calcquality_and_print_comparison(native_lib, dtype, capsys):
    """statistics.py:100-229 (the "next" row of SURVEY section 8f)."""
    import oracle
    from rfi_toolbox_b200.evaluation import compute_calcquality, print_statistics_comparison
    data, mask = make_cube(dtype=dtype, seed=71)
    noisy = mask ^ (np.random.default_rng(5).random(mask.shape) < 0.01)
    for kwargs in ({}, {"reference_data": data * dtype(0.5)}):
        got, want = compute_calcquality(data, noisy, **kwargs), oracle.compute_calcquality(data, noisy, **kwargs)
        assert set(got) == set(want)
        for k, v in want.items():
            if k == "components":
                for kk, vv in v.items():
                    assert got[k][kk] == pytest.approx(vv, rel=2e-6, abs=1e-9), kk
            else:
                assert got[k] == pytest.approx(v, rel=1e-5, abs=1e-6), k
    allf = compute_calcquality(data, np.ones_like(mask))
    assert allf["calcquality"] == np.inf and allf["components"] == {} and allf["flagged_pct"] == 100.0
    print_statistics_comparison(data, noisy)
    out = capsys.readouterr().out
    assert "Statistics Comparison (Before/After Flagging)" in out and "Flagging Fidelity Index (FFI):" in out
    st = oracle.compute_statistics(data, noisy)
    assert f"  Count:  {st['count']}" in out and f"({st['flagged_fraction']*100:.2f}% flagged)" in out


@pytest.mark.parametrize("dtype", [np.complex64, np.float32])
@pytest.mark.parametrize("n", [1, 5, 4099, 300_001, 5_000_003])
def test_whole_cube_statistics_three_pass_exact(native_lib, dtype, n):
    """`rfi_statistics2` (sampled brackets, one to three list levels depending on n): medians and MADs are
    exact order statistics of all samples and of the unflagged ones -- bit-identical to NumPy's -- for
    sizes below, at and far above one CTA's final list, odd and even counts, ragged tails."""
    import oracle
    from rfi_toolbox_b200 import compute_ffi, compute_statistics
    rng = np.random.default_rng(n)
    data = rng.normal(0, 1, n) + 1j * rng.normal(0, 1, n)
    flags = rng.random(n) < 0.1
    data[flags] *= 50.0
    flags ^= rng.random(n) < 0.01
    data = (data if np.dtype(dtype).kind == "c" else np.abs(data) * np.where(rng.random(n) < 0.3, -1.0, 1.0)).astype(dtype)
    for fl in (None, flags):
        got, want = compute_statistics(data, fl), oracle.compute_statistics(data, fl)
        for k in ("median", "mad"):
            assert (np.isnan(got[k]) and np.isnan(want[k])) or np.float32(got[k]) == np.float32(want[k]), (k, got[k], want[k])
        scale = float(np.mean(np.abs(data)))
        for k in ("mean", "std"):
            assert got[k] == pytest.approx(want[k], rel=1e-6, abs=2e-6 * scale, nan_ok=True), (k, got[k], want[k])
        assert got["count"] == want["count"] and got["flagged_fraction"] == pytest.approx(want["flagged_fraction"])
    if n > 5 and flags.any() and not flags.all():
        got, want = compute_ffi(data, flags), oracle.compute_ffi(data, flags)
        for k, v in want.items():
            assert got[k] == pytest.approx(v, rel=1e-5, abs=1e-7), (k, got[k], v)


def test_whole_cube_statistics_special_values(native_lib):
    import oracle
    from rfi_toolbox_b200 import compute_statistics
    rng = np.random.default_rng(3)
    base = np.abs(rng.normal(0, 1, 400_000)).astype(np.float32)
    flags = rng.random(base.size) < 0.2
    for name in ("nan_flagged", "nan_unflagged", "inf", "all_flagged", "duplicates"):
        d, f = base.copy(), flags.copy()
        if name == "nan_flagged":
            d[7] = np.nan; f[7] = True
        elif name == "nan_unflagged":
            d[7] = np.nan; f[7] = False
        elif name == "inf":
            d[11] = np.inf; d[12] = -np.inf
        elif name == "all_flagged":
            f[:] = True
        else:
            d = np.round(d * 4) / 4
        for fl in (None, f):
            got, want = compute_statistics(d, fl), oracle.compute_statistics(d, fl)
            for k in ("median", "mad", "mean", "std"):
                a, b = got[k], want[k]
                assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=2e-6, abs=1e-6) or (np.isinf(a) and a == b), (name, k, a, b)
            for k in ("median", "mad"):
                a, b = got[k], want[k]
                assert (np.isnan(a) and np.isnan(b)) or np.float32(a) == np.float32(b), (name, k, a, b)
