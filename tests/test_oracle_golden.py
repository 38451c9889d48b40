"""The CPU oracle against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py) and against the reference's own known-answer
tests for this path (tests/test_preprocessing.py:14-66, restated on oracle.tile)."""
import numpy as np
import pytest

import oracle
from tests.golden_util import CASE_NAMES, HOST_DEPENDENT_LABELS, load_case


@pytest.mark.parametrize("name", CASE_NAMES)
def test_create_dataset_matches_golden(name):
    c = load_case(name)
    np.random.seed(c["perm_seed"])
    ds = oracle.create_dataset(c["data"], c["flags"], **c["kwargs"])
    assert ds.labels.shape == c["labels"].shape
    if name in HOST_DEPENDENT_LABELS:
        assert (ds.labels != c["labels"]).mean() < 1e-4
    else:
        assert np.array_equal(ds.labels, c["labels"])
    got = ds.images.reshape(-1)[c["image_pos"]]
    assert np.allclose(got, c["image_val"], rtol=1e-6, atol=2e-5, equal_nan=True)
    assert int(np.isnan(ds.images).sum()) == c["image_nan_count"]
    assert np.allclose(np.nansum(ds.images.astype(np.float64), axis=(0, 1, 2)), c["image_channel_sum"],
                       rtol=1e-5, atol=1.0)
    meta = c["info"]["metadata"]
    assert ds.metadata["patch_size"] == meta["patch_size"] and ds.metadata["stretch"] == meta["stretch"]
    assert [list(s) for s in ds.metadata["original_shapes"]] == meta["original_shapes"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_metrics_and_ffi_match_golden(name):
    c = load_case(name)
    pred = c["labels"].astype(bool)
    truth = pred ^ (np.random.default_rng(9).random(pred.shape) < 0.02)
    ev = oracle.evaluate_segmentation(pred, truth)
    for k, v in c["evaluation"].items():
        assert float(ev[k]) == v
    ffi = oracle.compute_ffi(c["cube"], c["mask"])
    st = oracle.compute_statistics(c["cube"], c["mask"])
    for k, v in c["ffi"].items():
        assert ffi[k] == pytest.approx(v, rel=1e-6, abs=1e-12, nan_ok=True)  # |z| is host-SIMD dependent
    for k, v in c["stats"].items():
        assert float(st[k]) == pytest.approx(v, rel=1e-6, nan_ok=True)


class TestTileKnownAnswers:
    """rfi_toolbox tests/test_preprocessing.py:14-66 (patchify), on the oracle's tiler."""

    def test_shape_4x4(self):
        assert oracle.tile(np.arange(16).reshape(4, 4), 2).shape == (4, 2, 2)

    def test_content(self):
        t = oracle.tile(np.arange(16).reshape(4, 4), 2)
        assert np.array_equal(t[0], [[0, 1], [4, 5]]) and np.array_equal(t[-1], [[10, 11], [14, 15]])

    def test_large(self):
        assert oracle.tile(np.random.rand(1024, 1024), 128).shape == (64, 128, 128)

    def test_non_square(self):
        assert oracle.tile(np.arange(24).reshape(6, 4), 2).shape == (6, 2, 2)

    def test_single(self):
        a = np.arange(4).reshape(2, 2)
        t = oracle.tile(a, 2)
        assert t.shape == (1, 2, 2) and np.array_equal(t[0], a)

    def test_dtype(self):
        assert oracle.tile(np.array([[1.5, 2.5], [3.5, 4.5]], dtype=np.float32), 2).dtype == np.float32

    def test_padding(self):
        t = oracle.tile(np.ones((3, 5)), 2)  # preprocessor.py:527-550: zero-pad bottom / right
        assert t.shape == (6, 2, 2) and t.sum() == 15


def test_cabs_formula_is_numpy_abs():
    """The magnitude formula the CUDA kernels use (rfi_common.cuh cabs_np) == np.abs(complex64)."""
    rng = np.random.default_rng(0)
    for scale in (1.0, 1e6, 1e-20, 1e18):
        z = ((rng.standard_normal(200000) + 1j * rng.standard_normal(200000)) * scale).astype(np.complex64)
        assert np.array_equal(oracle.cabs_numpy_algorithm(z), np.abs(z))


def test_canonical_index_map_is_reference_order():
    """position of every rotated tile in the reference's patch list."""
    B, npol, C, T, P = 2, 2, 256, 384, 128
    ids = np.arange(B * npol * C * T, dtype=np.float64).reshape(B, npol, C, T)
    views = oracle.ref_port.rotate_views(ids, 4, True)
    patches, _ = oracle.ref_port.tile_all(views, P)
    cmap = oracle.canonical_index_map(B * npol, 4, C // P, T // P)
    for w in range(B * npol):
        wf = ids[w // npol, w % npol]
        for i in range(C // P):
            for j in range(T // P):
                x = wf[i * P:(i + 1) * P, j * P:(j + 1) * P]
                assert np.array_equal(patches[cmap[w, 0, i, j]], x)
                assert np.array_equal(patches[cmap[w, 1, i, j]], x[::-1, :])
                assert np.array_equal(patches[cmap[w, 2, i, j]], x.T)
                assert np.array_equal(patches[cmap[w, 3, i, j]], x.T[::-1, :])
