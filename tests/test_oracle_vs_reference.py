"""Pin the oracle to the LIVE reference (only where /root/reference exists, i.e. the build
container): bit-for-bit on the same host for every stage of the path."""
import os
import sys

import numpy as np
import pytest

import oracle
from tests.cubes import make_cube

pytestmark = pytest.mark.reference
os.environ.setdefault("CI", "1")  # reference skips its multiprocessing pools (preprocessor.py:491)


@pytest.fixture(scope="module")
def ref():
    sys.path.append("/root/reference")  # appended: this repo's `tests` package stays first
    from rfi_toolbox.evaluation import compute_ffi, compute_statistics, evaluate_segmentation
    from rfi_toolbox.preprocessing import Preprocessor
    from rfi_toolbox.preprocessing.preprocessor import _compute_mad_flag_single_patch, patchify
    return dict(Preprocessor=Preprocessor, evaluate_segmentation=evaluate_segmentation,
                compute_ffi=compute_ffi, compute_statistics=compute_statistics,
                mad_single=_compute_mad_flag_single_patch, patchify=patchify)


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


CASES = [
    (np.float32, dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)),
    (np.float32, dict(stretch="LOG10", flag_sigma=3.5, use_custom_flags=False, augmentation_rotations=2)),
    (np.float64, dict(stretch=None, flag_sigma=5, use_custom_flags=False, normalize_after_stretch=True)),
    (np.float32, dict(stretch="SQRT", use_custom_flags=True, num_patches=11)),
    (np.complex64, dict(use_custom_flags=True)),
    (np.complex128, dict(use_custom_flags=True, enable_augmentation=False)),
    (np.float32, dict(stretch="SQRT", inference_mode=True)),
]


@pytest.mark.parametrize("dtype,kw", CASES)
def test_create_dataset_bit_identical(ref, dtype, kw):
    data, mask = make_cube(dtype=dtype, seed=41, special=(dtype == np.float32))
    flags = mask if kw.get("use_custom_flags", True) and not kw.get("inference_mode") else None
    np.random.seed(5)
    r = ref["Preprocessor"](data, flags).create_dataset(patch_size=128, num_workers=0, **kw)
    np.random.seed(5)
    o = oracle.create_dataset(data, flags, patch_size=128, **kw)
    assert _same(r.images.numpy(), o.images) and _same(r.labels.numpy(), o.labels)
    assert r.metadata == o.metadata


def test_padding_and_skip_paths(ref):
    data, mask = make_cube(dtype=np.float32, seed=43)
    d2, m2 = data[:, :, :200, :300], mask[:, :, :200, :300]
    np.random.seed(6)
    r = ref["Preprocessor"](d2, m2).create_dataset(patch_size=128, num_workers=0)
    np.random.seed(6)
    o = oracle.create_dataset(d2, m2, patch_size=128)
    assert _same(r.images.numpy(), o.images) and _same(r.labels.numpy(), o.labels) and r.metadata == o.metadata
    d3 = data[:, :, :128, :128]
    r = ref["Preprocessor"](d3).create_dataset(patch_size=128, num_workers=0, inference_mode=True)
    o = oracle.create_dataset(d3, None, patch_size=128, inference_mode=True)
    assert _same(r.images.numpy(), o.images) and r.metadata == o.metadata


def test_mad_flags_complex_use_pool_semantics(ref):
    data, _ = make_cube(dtype=np.complex64, seed=45)
    tiles = oracle.tile(data[0, 0], 128)
    a = np.array([ref["mad_single"](t, 4) for t in tiles])
    assert _same(a, oracle.mad_flags(tiles, 4))


def test_metrics_and_statistics(ref):
    rng = np.random.default_rng(3)
    for shape in ((64, 64), (3, 128, 128)):
        p, t = rng.random(shape) < 0.3, rng.random(shape) < 0.3
        z = np.zeros(shape)
        for pp, tt in ((p, t), (p.astype(np.float32), t.astype(np.uint8)), (z, z), (z, t), (p, z)):
            a, b = ref["evaluate_segmentation"](pp, tt), oracle.evaluate_segmentation(pp, tt)
            assert a == b and all(type(a[k]) is type(b[k]) for k in a)
    for dtype in (np.float32, np.float64, np.complex64, np.complex128):
        data, mask = make_cube(dtype=dtype, seed=47)
        assert ref["compute_ffi"](data, mask) == oracle.compute_ffi(data, mask)
        assert ref["compute_statistics"](data, mask) == oracle.compute_statistics(data, mask)
        assert ref["compute_statistics"](data) == oracle.compute_statistics(data)
        assert ref["compute_ffi"](data, np.ones_like(mask)) == oracle.compute_ffi(data, np.ones_like(mask))
        from rfi_toolbox.evaluation.statistics import compute_calcquality
        noisy = mask ^ (rng.random(mask.shape) < 0.01)
        assert compute_calcquality(data, noisy) == oracle.compute_calcquality(data, noisy)
        assert compute_calcquality(data, noisy, reference_data=data * 0.5) == \
            oracle.compute_calcquality(data, noisy, reference_data=data * 0.5)
        assert compute_calcquality(data, np.ones_like(mask)) == oracle.compute_calcquality(data, np.ones_like(mask))


def test_patchify_and_tile_agree(ref):
    a = np.arange(24 * 36, dtype=np.float32).reshape(24, 36)
    assert _same(ref["patchify"](a, (12, 12), 12).reshape(-1, 12, 12), oracle.tile(a, 12))
    from rfi_toolbox_b200.preprocessing import patchify
    assert _same(ref["patchify"](a, (12, 12), 12), patchify(a, (12, 12), 12))
