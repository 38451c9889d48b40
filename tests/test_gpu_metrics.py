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


def test_sqrt_unit_range_is_exact(native_lib):
    """cabs_fast's square root (no range test, argument in [1, 2]) equals sqrt.rn.f32 on every
    one of the 2^23 + 1 possible arguments."""
    import torch
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = native_lib.rfi_selftest_sqrt_unit(bad.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert int(bad.item()) == 0
