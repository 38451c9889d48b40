"""Single-launch path (`rfi_fused_patches`, tile_fused_kernel): statistics and patches of a tile in ONE
kernel, destination slots drawn ahead (speculating that no patch is blank).

Checked here: the single launch gives bit-identical results to the two-launch path and to the oracle
(labels, patch order) within the image noise bound; the global NumPy stream ends where the reference
leaves it -- also when the speculation FAILS (a blank tile among flagged ones), alone or in the middle
of several calls in flight; tiles that leave the monotone algorithm inside the fused CTA (NaN, inf,
negative samples, LOG10 of exact zeros) still match.
"""
import numpy as np
import pytest
import torch

import oracle
from tests.cubes import make_cube
from tests.test_gpu_parity import IMG_ATOL, IMG_RTOL, _compare, _run_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fresh_speculation_state():
    """The cool-down of call signatures that mis-speculated is process-wide: start every test without it."""
    from rfi_toolbox_b200.preprocessing import preprocessor as mod
    for d in (mod._SPEC_COOLDOWN, mod._SPEC_BACKOFF):
        d.clear()
    mod.Preprocessor.speculate = True     # opt-in (off by default: the two-launch path is faster)
    yield
    mod.Preprocessor.speculate = False
    for d in (mod._SPEC_COOLDOWN, mod._SPEC_BACKOFF):
        d.clear()

CASES = [
    dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False),
    dict(stretch="SQRT", flag_sigma=3.5, use_custom_flags=False, augmentation_rotations=2),
    dict(stretch=None, flag_sigma=5, use_custom_flags=False, enable_augmentation=False),
    dict(stretch="SQRT", flag_sigma=4, use_custom_flags=False, normalize_after_stretch=True),
    dict(stretch=None, flag_sigma=5, use_custom_flags=False, normalize_before_stretch=False, num_patches=13),
]


def _run(data, speculate, magnitude, seed=11, **kw):
    from rfi_toolbox_b200 import Preprocessor
    np.random.seed(seed)
    pre = Preprocessor(data, None, magnitude=magnitude)
    pre.speculate = speculate
    ds = pre.create_dataset(**kw)
    torch.cuda.synchronize()
    return pre, ds, np.random.random()


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_single_launch_equals_two_launches_and_oracle(native_lib, kw, dtype):
    data, _ = make_cube(dtype=dtype, seed=3)
    mag = dtype == np.complex64
    pre1, ds1, after1 = _run(data, True, mag, **kw)
    pre2, ds2, after2 = _run(data, False, mag, **kw)
    assert pre1.last_launch == "single" and pre2.last_launch == "two-phase"
    assert after1 == after2, "the global NumPy stream ends elsewhere"
    assert np.array_equal(pre1.order, pre2.order)
    assert torch.equal(ds1.labels, ds2.labels)
    assert torch.equal(ds1.images, ds2.images), \
        f"max |single - two| = {(ds1.images - ds2.images).abs().max().item():.3e}"
    ods, inter = _run_oracle(data, None, magnitude=mag, **kw)
    _compare(ds1, ods, inter, pre1, label=f"single-launch {np.dtype(dtype).name} {kw.get('stretch')}")
    # the statistics the single launch leaves are the two-launch ones
    from rfi_toolbox_b200 import _native
    s1 = np.frombuffer(pre1.last_tile_stats.cpu().numpy().tobytes(), dtype=np.uint8).reshape(-1, _native.TILE_STAT_BYTES)
    s2 = np.frombuffer(pre2.last_tile_stats.cpu().numpy().tobytes(), dtype=np.uint8).reshape(-1, _native.TILE_STAT_BYTES)
    assert np.array_equal(s1, s2)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_single_launch_special_values(native_lib, dtype):
    """NaN / inf / exact zero samples: those tiles leave the monotone algorithm INSIDE the fused CTA
    (general statistics, pass A from the cube) and must still match."""
    data, _ = make_cube(dtype=dtype, seed=5, special=True)
    if dtype == np.float32:
        data[1, 0, 140:150, 20:40] *= -1.0  # negative samples in one tile
    mag = dtype == np.complex64
    kw = dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    pre, ds, _ = _run(data, True, mag, **kw)
    assert pre.last_launch == "single"
    ods, inter = _run_oracle(data, None, magnitude=mag, **kw)
    _compare(ds, ods, inter, pre, label="single-launch special values")


def test_single_launch_log10_zero_rows(native_lib):
    data, _ = make_cube(dtype=np.complex64, seed=7)
    kw = dict(stretch="LOG10", flag_sigma=5, use_custom_flags=False)
    pre, ds, _ = _run(data, True, True, **kw)
    assert pre.last_launch == "single"
    pre2, ds2, _ = _run(data, False, True, **kw)
    assert torch.equal(ds.labels, ds2.labels)
    # the tiles with exact-zero rows go through the general algorithm on both sides, which also finds their raw
    # thresholds (route GENERAL | RAW_THRESHOLDS | RAW_FILL = 11) so that phase 2 takes its fast route
    from tests.test_gpu_bigtile import _routes
    s1, s2 = _routes(pre), _routes(pre2)
    fill = (s2["route"] & 8) != 0
    assert fill.sum() >= 4 and ((s2["route"][fill] & 11) == 11).all(), s2["route"]
    for k in ("median_before", "inf_fill", "centre", "mad", "thr_lo", "thr_hi", "n_valid", "n_inf", "n_flagged", "raw_lo", "raw_hi"):
        assert np.array_equal(s1[k], s2[k]), (k, s1[k][fill], s2[k][fill])
    assert torch.equal(ds.images, ds2.images)
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre, exact_labels=False, max_label_mismatch=1e-4, label="single-launch LOG10")
    _compare(ds2, ods, inter, pre2, exact_labels=False, max_label_mismatch=1e-4, label="two-launch LOG10")


def test_single_launch_inference_mode(native_lib):
    from rfi_toolbox_b200 import Preprocessor
    data, mask = make_cube(dtype=np.float32, seed=9)
    kw = dict(stretch="SQRT", flag_sigma=5, inference_mode=True)
    np.random.seed(4)
    pre = Preprocessor(data, mask)
    ds = pre.create_dataset(**kw)
    assert pre.last_launch == "single"
    after = np.random.random()
    np.random.seed(4)
    ods, inter = oracle.create_dataset(data, mask, return_intermediates=True, **kw)
    assert after == np.random.random()
    _compare(ds, ods, inter, pre, label="single-launch inference")


def _cube_with_blank_tiles(seed, dtype=np.float32):
    """Two tiles of constant samples: MAD = 0, thresholds = the constant itself, no sample flagged ->
    blank patches among flagged ones: the all-kept speculation is wrong for this cube."""
    data, _ = make_cube(n_bl=1, n_pol=2, dtype=dtype, seed=seed)
    data[0, 0, 0:128, 128:256] = 3.0
    data[0, 1, 128:256, 0:128] = 0.5
    return data


def test_failed_speculation_completes_through_two_phases(native_lib):
    kw = dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    data = _cube_with_blank_tiles(61)
    pre, ds, after = _run(data, True, False, seed=13, **kw)
    assert pre.last_launch.startswith("single, then phase 2")
    np.random.seed(13)
    ods, inter = oracle.create_dataset(data, None, return_intermediates=True, **kw)
    assert after == np.random.random(), "the global NumPy stream ends elsewhere after a failed speculation"
    assert len(ds) == len(ods) < 2 * 4 * 6
    _compare(ds, ods, inter, pre, label="failed speculation")
    # the signature now cools down: the next call takes two launches straight away
    pre2, ds2, _ = _run(data, True, False, seed=13, **kw)
    assert pre2.last_launch == "two-phase"
    assert torch.equal(ds2.labels, ds.labels) and torch.equal(ds2.images, ds.images)


def test_failed_speculation_with_calls_in_flight(native_lib):
    """Four calls in flight, the second one mis-speculates: every call's permutation must still be the one
    sequential reference calls draw, and the stream must end where theirs does."""
    from rfi_toolbox_b200 import Preprocessor
    kw = dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    cubes = [make_cube(n_bl=1, n_pol=2, dtype=np.float32, seed=70)[0], _cube_with_blank_tiles(71),
             make_cube(n_bl=1, n_pol=2, dtype=np.float32, seed=72)[0], make_cube(n_bl=1, n_pol=2, dtype=np.float32, seed=73)[0]]
    np.random.seed(33)
    want = [oracle.create_dataset(c, None, return_intermediates=True, **kw) for c in cubes]
    after_ref = np.random.random()
    np.random.seed(33)
    pres = [Preprocessor(c, None) for c in cubes]
    pend = [p.create_dataset_async(**kw) for p in pres[:3]]       # three in flight, the middle one wrong
    out = [pend[0].result(), pend[1].result()]
    pend.append(pres[3].create_dataset_async(**kw))                # submitted while call 2 awaits its re-draw
    out += [pend[2].result(), pend[3].result()]
    assert np.random.random() == after_ref
    kinds = [p.last_launch for p in pres]
    assert kinds[0] == "single" and kinds[1].startswith("single, then") and kinds[2].startswith("single, then")
    assert kinds[3] == "two-phase"
    for p, ds, (ods, inter) in zip(pres, out, want):
        assert np.array_equal(p.order, inter["order"])
        assert np.array_equal(ds.labels.cpu().numpy(), ods.labels)
        assert np.allclose(ds.images.cpu().numpy(), ods.images, rtol=IMG_RTOL, atol=IMG_ATOL, equal_nan=True)


@pytest.mark.parametrize("chunks", [1, 3])
def test_single_launch_pinned_host_input(native_lib, chunks):
    from rfi_toolbox_b200 import Preprocessor
    data, _ = make_cube(n_bl=3, n_pol=2, dtype=np.complex64, seed=81)
    kw = dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    np.random.seed(11)
    pre = Preprocessor(torch.from_numpy(data).pin_memory(), None, magnitude=True, pin=True)
    pre.upload_chunks = chunks
    ds = pre.create_dataset(**kw)
    torch.cuda.synchronize()
    assert pre.last_launch == "single"
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre, label="single-launch pinned host input")
