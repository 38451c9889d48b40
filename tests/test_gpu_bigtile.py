"""GPU parity of the BIG-TILE create_dataset path (csrc/rfi_bigtile.cu): P = 256 / 512 / 1024 with
dims that are multiples of P, float32 arithmetic, real branch -- BASELINE config 5's geometry.
The oracle is the CPU restatement of the reference; bars as in tests/test_gpu_parity.py
(labels, patch order, statistics bit-exact; images within 1e-6 relative + 2e-5)."""
import numpy as np
import pytest
import torch

from tests.cubes import make_cube
from tests.test_gpu_parity import _compare, _run_gpu, _run_oracle

pytestmark = pytest.mark.gpu


def _routes(pre):
    from rfi_toolbox_b200 import _native
    raw = pre.last_tile_stats.cpu().numpy()
    st = np.frombuffer(raw.tobytes(), dtype=np.dtype([
        ("median_before", "f8"), ("inf_fill", "f8"), ("median_after", "f8"), ("centre", "f8"), ("mad", "f8"),
        ("thr_lo", "f8"), ("thr_hi", "f8"), ("n_valid", "i4"), ("n_inf", "i4"), ("n_flagged", "i4"),
        ("route", "i4"), ("raw_lo", "f8"), ("raw_hi", "f8")]))
    assert st.itemsize == _native.TILE_STAT_BYTES
    return st


CASES = [
    dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False),
    dict(stretch=None, flag_sigma=3, use_custom_flags=False),                     # config 5's call
    dict(stretch="SQRT", flag_sigma=3.5, use_custom_flags=False, augmentation_rotations=2),
    dict(stretch=None, flag_sigma=5, use_custom_flags=False, enable_augmentation=False),
    dict(stretch="SQRT", flag_sigma=4, use_custom_flags=False, normalize_after_stretch=True),
    dict(stretch=None, flag_sigma=5, use_custom_flags=False, normalize_before_stretch=False, num_patches=5),
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.complex64])
def test_p256_mad_flags_bit_exact(native_lib, kw, dtype):
    data, _ = make_cube(n_bl=2, n_pol=2, channels=512, times=768, dtype=dtype, seed=71)
    mag = dtype == np.complex64
    pre, ds = _run_gpu(data, None, magnitude=mag, patch_size=256, **kw)
    ods, inter = _run_oracle(data, None, magnitude=mag, patch_size=256, **kw)
    _compare(ds, ods, inter, pre)
    st = _routes(pre)
    # the sampled brackets settle (nearly) every group: these cubes have no special values
    assert (st["route"] & 1).mean() > 0.9


@pytest.mark.parametrize("patch,shape", [(512, (1024, 1536)), (1024, (1024, 2048))])
def test_p512_p1024(native_lib, patch, shape):
    data, _ = make_cube(n_bl=1, n_pol=2, channels=shape[0], times=shape[1], dtype=np.float32, seed=73)
    kw = dict(patch_size=patch, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)
    assert (_routes(pre)["route"] & 1).all()


def test_p256_special_values_fall_back(native_lib):
    """NaN / +inf / exact zero samples: those groups are measured by the generic select
    (route GENERAL), the others keep the raw thresholds; results identical either way."""
    data, _ = make_cube(n_bl=2, n_pol=2, channels=512, times=768, dtype=np.float32, seed=75, special=True)
    kw = dict(patch_size=256, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)
    st = _routes(pre)
    assert (st["route"] & 2).sum() >= 2 and (st["route"] & 1).sum() >= 1


def test_p256_log10_zero_rows(native_lib):
    """LOG10: the exact-zero bandpass edge rows give -inf -> MAD fill -> generic select for the
    groups of the first / last tile row; float32 flags sit downstream of a non-reproducible log10."""
    data, _ = make_cube(n_bl=1, n_pol=2, channels=1024, times=512, dtype=np.float32, seed=77)
    kw = dict(patch_size=256, stretch="LOG10", flag_sigma=5, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre, exact_labels=False, max_label_mismatch=1e-4)
    st = _routes(pre)
    assert (st["route"] & 2).any() and (st["route"] & 1).any()


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("rot", [1, 4])
def test_p256_custom_flags(native_lib, norm, rot):
    """custom flags: with normalisation (median only, no MAD pass) and without (no statistic)."""
    data, mask = make_cube(n_bl=2, n_pol=2, channels=512, times=512, dtype=np.float32, seed=79)
    kw = dict(patch_size=256, stretch="SQRT" if norm else None, use_custom_flags=True,
              normalize_before_stretch=norm, augmentation_rotations=rot)
    pre, ds = _run_gpu(data, mask, **kw)
    ods, inter = _run_oracle(data, mask, **kw)
    _compare(ds, ods, inter, pre)


def test_p256_inference_mode(native_lib):
    data, _ = make_cube(n_bl=1, n_pol=2, channels=512, times=512, dtype=np.complex64, seed=81)
    kw = dict(patch_size=256, stretch="SQRT", inference_mode=True)
    pre, ds = _run_gpu(data, None, magnitude=True, **kw)
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre)
    assert int(ds.labels.sum()) == 0


def test_chunked_stream_matches_per_chunk_oracle(native_lib):
    """iter_dataset_chunks: each baseline chunk is one Preprocessor (own blank removal, own
    shuffle drawn in chunk order from the global generator) -- config 5's streaming mode."""
    import oracle
    from rfi_toolbox_b200.preprocessing import iter_dataset_chunks
    data, _ = make_cube(n_bl=3, n_pol=2, channels=256, times=512, dtype=np.float32, seed=91)
    kw = dict(patch_size=256, stretch=None, flag_sigma=3, use_custom_flags=False)
    np.random.seed(5)
    got = [(b0, b1, ds.images.cpu().numpy(), ds.labels.cpu().numpy())
           for b0, b1, ds in iter_dataset_chunks(data, chunk_baselines=2, **kw)]
    assert [(g[0], g[1]) for g in got] == [(0, 2), (2, 3)]
    np.random.seed(5)
    for b0, b1, imgs, labs in got:
        ods = oracle.create_dataset(data[b0:b1], None, **kw)
        assert np.array_equal(labs, ods.labels)
        assert np.allclose(imgs, ods.images, rtol=1e-6, atol=3e-6, equal_nan=True)


def _degenerate_cube(seed=0):
    """Tiles that stress the sampled brackets: heavy duplicates, two-valued, constant, all-zero."""
    rng = np.random.default_rng(seed)
    d = np.abs(rng.normal(1.0, 0.1, (2, 2, 512, 512))).astype(np.float32)
    d[0, 0] = np.round(d[0, 0] * 8) / 8                       # ~10 distinct values
    d[0, 1] = np.where(rng.random((512, 512)) < 0.5, 1.0, 2.0)  # two values: the median sits on a tie
    d[1, 0, :256] = 3.0                                       # constant tiles (MAD 0)
    d[1, 0, 256:, :256] = 0.0                                 # all-zero tile
    d[1, 1, 100:110, :] = 1e6                                 # a clean RFI line in plain noise
    return d


@pytest.mark.parametrize("patch", [128, 256])
@pytest.mark.parametrize("stretch", [None, "SQRT"])
def test_degenerate_tiles_match_oracle(native_lib, patch, stretch):
    """Ties everywhere: exact order statistics must still come out bit-identical (duplicates end a
    histogram refinement on a single key value; constant tiles have MAD 0 and flag nothing)."""
    data = _degenerate_cube()
    kw = dict(patch_size=patch, stretch=stretch, flag_sigma=3, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)


@pytest.mark.parametrize("rot", [2, 4])
def test_p256_padded_views(native_lib, rot):
    """Dims that are not multiples of 256: one single-view big-tile plan per rotation over rotated,
    zero-padded copies; every rotated patch is its own statistics group (preprocessor.py:527-550)."""
    data, _ = make_cube(n_bl=1, n_pol=2, channels=300, times=600, dtype=np.complex64, seed=93)
    kw = dict(patch_size=256, stretch="SQRT", flag_sigma=4, use_custom_flags=False, augmentation_rotations=rot)
    pre, ds = _run_gpu(data, None, magnitude=True, **kw)
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre)
    assert (_routes(pre)["route"] & 3).all()


@pytest.mark.parametrize("patch,shape", [(256, (512, 768)), (512, (1024, 512)), (1024, (1024, 1024))])
@pytest.mark.parametrize("rot", [1, 4])
def test_complex_branch_on_chip(native_lib, patch, shape, rot):
    """Complex input + custom flags at P >= 256 (gradient / log-amplitude / phase channels,
    preprocessor.py:562-606) through the big-tile kernels.  The last case is the reference's own production
    call: 1024 x 1024 complex waterfalls at patch_size = 1024 (configs/data_generation/synthetic_train_4k.yaml:50,
    synthetic_generator.py:93-105), where patchify is skipped (:261) and the patches are the rotated waterfalls."""
    from rfi_toolbox_b200 import _native
    data, mask = make_cube(n_bl=1, n_pol=2, channels=shape[0], times=shape[1], dtype=np.complex64, seed=81)
    kw = dict(patch_size=patch, use_custom_flags=True, augmentation_rotations=rot, enable_augmentation=rot > 1)
    pre, ds = _run_gpu(data, mask, **kw)
    ods, inter = _run_oracle(data, mask, **kw)
    _compare(ds, ods, inter, pre, label=f"complex branch P={patch} R={rot}")
    import ctypes as C
    from rfi_toolbox_b200.preprocessing.preprocessor import _DTYPE_CODE  # noqa: F401
    plan = _native.RfiPlan(dtype=_native.RFI_C64, magnitude=0, n_waterfalls=2, channels=shape[0], times=shape[1], patch=patch,
                           rotations=rot, stretch=0, norm_before=0, norm_after=0, flag_mode=_native.RFI_FLAGS_CUSTOM, sigma=5.0)
    assert native_lib.rfi_plan_path(C.byref(plan)) == _native.RFI_PATH_BIG


def test_complex_branch_on_chip_inference_and_patches(native_lib):
    data, mask = make_cube(n_bl=1, n_pol=2, channels=512, times=512, dtype=np.complex64, seed=82)
    kw = dict(patch_size=256, inference_mode=True)
    pre, ds = _run_gpu(data, mask, **kw)
    ods, inter = _run_oracle(data, mask, **kw)
    _compare(ds, ods, inter, pre, label="complex branch P=256 inference")
    assert np.array_equal(pre.patches.cpu().numpy(), inter["processed"][inter["order"]])


@pytest.mark.parametrize("patch", [256, 512])
def test_skipped_patchify_square_waterfall_on_chip(native_lib, patch):
    """Real branch, waterfall of exactly P x P: the reference skips patchify (preprocessor.py:261); one tile
    per waterfall through the big-tile kernels gives the same R rotated patches."""
    data, _ = make_cube(n_bl=2, n_pol=2, channels=patch, times=patch, dtype=np.float32, seed=83)
    kw = dict(patch_size=patch, stretch="SQRT", flag_sigma=4, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre, label=f"skipped patchify P={patch}")
    assert ds.metadata["original_shapes"] is None
