"""GPU parity of the GENERIC create_dataset path (csrc/rfi_generic.cu) against the CPU oracle:
patch sizes other than 128, dims that are not multiples of the patch size (the reference pads
after the rotation, preprocessor.py:527-550) and waterfalls no larger than the patch
(patchify skipped, preprocessor.py:261-269).  Same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

from tests.cubes import make_cube
from tests.test_gpu_parity import _compare, _run_gpu, _run_oracle

pytestmark = pytest.mark.gpu

MAD_SQRT = dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)


@pytest.mark.parametrize("patch", [64, 256])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_other_patch_sizes_divisible(native_lib, patch, dtype):
    """dims multiples of P: the R rotations of a tile share one statistics group."""
    data, _ = make_cube(n_bl=2, n_pol=2, channels=512, times=768, dtype=dtype, seed=31)
    kw = dict(patch_size=patch, **MAD_SQRT)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)


def test_patch_512_and_1024(native_lib):
    """The two largest legal patch sizes (config/validators.py:26-39): 256 Ki and 1 Mi samples
    per median -- far beyond one CTA; many CTAs cooperate per group."""
    data, _ = make_cube(n_bl=1, n_pol=2, channels=1024, times=2048, dtype=np.float32, seed=37)
    for patch in (512, 1024):
        kw = dict(patch_size=patch, **MAD_SQRT)
        pre, ds = _run_gpu(data, None, **kw)
        ods, inter = _run_oracle(data, None, **kw)
        _compare(ds, ods, inter, pre)


@pytest.mark.parametrize("rot", [1, 2, 4])
@pytest.mark.parametrize("patch,shape", [(100, (256, 384)), (128, (200, 300)), (128, (128, 200)), (64, (50, 130))])
def test_padded_mad_flags(native_lib, rot, patch, shape):
    """Zero padding follows the flip / transpose, so every rotated patch has its own statistics."""
    data, _ = make_cube(n_bl=2, n_pol=2, channels=shape[0], times=shape[1], dtype=np.float32, seed=41)
    kw = dict(patch_size=patch, augmentation_rotations=rot, **MAD_SQRT)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)


@pytest.mark.parametrize("dtype", [np.float32, np.complex64, np.complex128])
def test_padded_custom_flags(native_lib, dtype):
    data, mask = make_cube(n_bl=2, n_pol=2, channels=200, times=300, dtype=dtype, seed=43)
    kw = dict(patch_size=128, stretch=None, use_custom_flags=True, normalize_before_stretch=False)
    pre, ds = _run_gpu(data, mask, **kw)
    ods, inter = _run_oracle(data, mask, **kw)
    _compare(ds, ods, inter, pre)


def test_padded_log10_inf_fill(native_lib):
    """LOG10 of the zero pad is -inf -> replaced by the MAD of the finite values of THAT patch
    (preprocessor.py:697-702); float64 keeps the flags reproducible."""
    data, _ = make_cube(n_bl=1, n_pol=2, channels=200, times=300, dtype=np.float64, seed=47, special=True)
    kw = dict(patch_size=128, stretch="LOG10", flag_sigma=4, use_custom_flags=False)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)


def test_normalize_after_and_magnitude_generic(native_lib):
    data, _ = make_cube(n_bl=1, n_pol=2, channels=256, times=384, dtype=np.complex64, seed=53)
    kw = dict(patch_size=64, stretch="SQRT", flag_sigma=4, use_custom_flags=False, normalize_after_stretch=True)
    pre, ds = _run_gpu(data, None, magnitude=True, **kw)
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre)


@pytest.mark.parametrize("rot,shape", [(2, (256, 384)), (4, (256, 256)), (1, (100, 60))])
def test_patchify_skipped(native_lib, rot, shape):
    """Waterfall no larger than the patch: every rotated waterfall is one patch."""
    data, mask = make_cube(n_bl=2, n_pol=2, channels=shape[0], times=shape[1], dtype=np.float32, seed=59)
    kw = dict(patch_size=512, augmentation_rotations=rot, **MAD_SQRT)
    pre, ds = _run_gpu(data, None, **kw)
    ods, inter = _run_oracle(data, None, **kw)
    _compare(ds, ods, inter, pre)
    kw = dict(patch_size=512, augmentation_rotations=rot, stretch=None, use_custom_flags=True)
    pre, ds = _run_gpu(data, mask, **kw)
    ods, inter = _run_oracle(data, mask, **kw)
    _compare(ds, ods, inter, pre)


def test_patchify_skipped_ragged_raises(native_lib):
    from rfi_toolbox_b200 import Preprocessor
    data, _ = make_cube(n_bl=1, n_pol=1, channels=100, times=60, dtype=np.float32, seed=61)
    with pytest.raises(ValueError):
        Preprocessor(data).create_dataset(patch_size=512, use_custom_flags=False)


def test_generic_matches_fast_path_statistics(native_lib):
    """A 256-patch is four 128-tiles: its median must lie between theirs, and the generic
    path's labels for P = 256 on a constant-statistics cube equal the fast path's."""
    from rfi_toolbox_b200 import Preprocessor
    rng = np.random.default_rng(5)
    data = np.abs(rng.normal(1.0, 0.1, (1, 1, 256, 256))).astype(np.float32)
    data[0, 0, 17, :] = 50.0
    np.random.seed(0)
    a = Preprocessor(data).create_dataset(patch_size=256, enable_augmentation=False, **MAD_SQRT)
    np.random.seed(0)
    b = Preprocessor(data).create_dataset(patch_size=128, enable_augmentation=False, **MAD_SQRT)
    torch.cuda.synchronize()
    assert a.labels.shape == (1, 256, 256) and b.labels.shape[1:] == (128, 128)
    assert int(a.labels[0, 17].sum()) == 256


def test_padded_p128_takes_the_on_chip_kernels(native_lib):
    """Dims that are not multiples of 128 run as one single-view plan per rotation over rotated,
    zero-padded copies (`rfi_rotate_pad`), i.e. through the P = 128 kernels: their tile statistics
    carry a route bit, the generic path's do not."""
    from tests.test_gpu_bigtile import _routes
    data, _ = make_cube(n_bl=2, n_pol=2, channels=200, times=300, dtype=np.complex64, seed=67)
    kw = dict(patch_size=128, **MAD_SQRT)
    pre, ds = _run_gpu(data, None, magnitude=True, **kw)
    ods, inter = _run_oracle(data, None, magnitude=True, **kw)
    _compare(ds, ods, inter, pre)
    st = _routes(pre)
    assert len(st) == 2 * 2 * 4 * 2 * 3 and (st["route"] & 3).all()
    # an odd patch size still takes the generic path
    pre2, _ = _run_gpu(data, None, magnitude=True, patch_size=100, **MAD_SQRT)
    assert not (_routes(pre2)["route"] & 3).any()
