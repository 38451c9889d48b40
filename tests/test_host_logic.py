"""Host-side logic of the drop-in API (no GPU): index maps, patchify known answers, metric
ratios, container behaviour, argument validation, and 'no CPU fallback'."""
import numpy as np
import pytest
import torch

import oracle
from rfi_toolbox_b200.datasets import TorchDataset
from rfi_toolbox_b200.evaluation.metrics import _ratios
from rfi_toolbox_b200.preprocessing import Preprocessor, canonical_index_map, patchify


class TestPatchify:
    """Same known answers as the reference's tests/test_preprocessing.py:14-66."""

    def test_basic_shape(self):
        assert patchify(np.arange(16).reshape(4, 4), (2, 2), step=2).shape == (2, 2, 2, 2)

    def test_content(self):
        p = patchify(np.arange(16).reshape(4, 4), (2, 2), step=2)
        np.testing.assert_array_equal(p[0, 0], [[0, 1], [4, 5]])
        np.testing.assert_array_equal(p[1, 1], [[10, 11], [14, 15]])

    def test_large(self):
        assert patchify(np.random.rand(1024, 1024), (128, 128), step=128).shape == (8, 8, 128, 128)

    def test_non_square(self):
        assert patchify(np.arange(24).reshape(6, 4), (2, 2), step=2).shape == (3, 2, 2, 2)

    def test_single(self):
        a = np.arange(4).reshape(2, 2)
        p = patchify(a, (2, 2), step=2)
        assert p.shape == (1, 1, 2, 2)
        np.testing.assert_array_equal(p[0, 0], a)

    def test_dtype(self):
        assert patchify(np.array([[1.5, 2.5], [3.5, 4.5]], dtype=np.float32), (2, 2), step=2).dtype == np.float32


@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (3, 4, 2, 3), (2, 2, 8, 1)])
@pytest.mark.parametrize("rot", [1, 2, 4])
def test_canonical_index_map(shape, rot):
    nwf, _, nh, nw = shape
    a = canonical_index_map(nwf, rot, nh, nw)
    assert np.array_equal(a, oracle.canonical_index_map(nwf, rot, nh, nw))
    assert sorted(a.ravel().tolist()) == list(range(nwf * rot * nh * nw))  # a permutation


def test_effective_rotations():
    f = Preprocessor._effective_rotations
    assert [f(True, r) for r in (0, 1, 2, 3, 4, 8)] == [1, 1, 2, 2, 4, 4]
    assert f(False, 4) == 1


def test_ratios_match_oracle_formulas():
    rng = np.random.default_rng(0)
    triples = [(0, 0, 0), (0, 5, 0), (0, 0, 7), (3, 0, 0), (1, 2, 3), (2**40, 2**41, 3)]
    triples += [tuple(int(v) for v in rng.integers(0, 1000, 3)) for _ in range(200)]
    for tp, fp, fn in triples:
        a, b = _ratios(tp, fp, fn), oracle.metrics_from_counts(tp, fp, fn)
        assert a == b and all(type(a[k]) is type(b[k]) for k in a)


def test_torch_dataset_contract(tmp_path):
    img = torch.zeros((3, 4, 4, 3), dtype=torch.float32)
    lab = torch.ones((3, 4, 4), dtype=torch.uint8)
    ds = TorchDataset(img, lab, {"patch_size": 4})
    assert len(ds) == 3 and set(ds[0]) == {"image", "label"} and ds[0]["image"].shape == (4, 4, 3)
    assert ds["data"].shape == (3, 3, 4, 4) and ds["labels"] is ds.labels and ds["metadata"]["patch_size"] == 4
    with pytest.raises(AssertionError):
        TorchDataset(img.double(), lab)
    with pytest.raises(AssertionError):
        TorchDataset(img, lab.int())
    with pytest.raises(AssertionError):
        TorchDataset(img[:2], lab)
    ds.save_to_disk(tmp_path / "d.pt")
    back = TorchDataset.load_from_disk(tmp_path / "d.pt")
    assert torch.equal(back.images, img) and back.metadata == {"patch_size": 4}


def test_preprocessor_argument_errors():
    with pytest.raises(ValueError):
        Preprocessor(np.zeros((4, 4)))  # ndim not in {3, 4} (preprocessor.py:191)
    p = Preprocessor(np.zeros((2, 128, 128), dtype=np.float32))
    assert p.data.shape == (1, 2, 128, 128)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    """Without a CUDA device every operator raises instead of silently computing on the host."""
    from rfi_toolbox_b200 import compute_ffi, evaluate_segmentation
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Preprocessor(np.ones((1, 1, 128, 128), dtype=np.float32)).create_dataset()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        evaluate_segmentation(np.zeros(8, bool), np.zeros(8, bool))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compute_ffi(np.ones(8, np.float32), np.zeros(8, bool))


@pytest.mark.parametrize("rot", [1, 2, 4])
def test_keep_mask_in_canonical_order(rot):
    from rfi_toolbox_b200.preprocessing.preprocessor import _keep_in_canonical_order
    rng = np.random.default_rng(4)
    kt = rng.random((5, 3, 7)) < 0.5
    cmap = canonical_index_map(5, rot, 3, 7)
    want = np.zeros(cmap.size, dtype=bool)
    want[cmap.ravel()] = np.broadcast_to(kt[:, None], cmap.shape).ravel()
    assert np.array_equal(_keep_in_canonical_order(kt, rot), want)


def test_legacy_permutation_matches_numpy(native_lib):
    """The native shuffle (csrc/rfi_host.cu) reproduces np.random.permutation of the global
    legacy generator -- values AND stream position -- so create_dataset consumes the RNG
    exactly like preprocessor.py:758-763."""
    import numpy as np
    from rfi_toolbox_b200 import _native
    for seed in range(6):
        # (16383 .. 32769 straddle the mask at which the vectorised draws switch block size)
        for n in [0, 1, 2, 3, 5, 8, 9, 33, 100, 623, 624, 625, 1248, 4097, 16383, 16384, 16385, 32767, 32769, 44368, 65537, 131073]:
            np.random.seed(seed); np.random.random(seed * 13)
            a = np.random.permutation(n); ra = np.random.random(5); ga = np.random.standard_normal(3)
            np.random.seed(seed); np.random.random(seed * 13)
            b = _native.legacy_permutation(n); rb = np.random.random(5); gb = np.random.standard_normal(3)
            assert a.dtype == b.dtype and np.array_equal(a, b), (seed, n)
            assert np.array_equal(ra, rb) and np.array_equal(ga, gb), (seed, n)


@pytest.mark.parametrize("shape,patch,rot", [((256, 384), 128, 4), ((256, 384), 128, 2), ((256, 384), 128, 1),
                                             ((200, 300), 128, 4), ((100, 60), 512, 1), ((256, 256), 512, 4),
                                             ((250, 330), 100, 4), ((512, 256), 256, 2)])
def test_plan_slots_matches_numpy_pipeline(native_lib, shape, patch, rot):
    """rfi_plan_slots == keep mask in canonical order -> np.random.permutation -> dest (the
    host phase of create_dataset written with NumPy), including the RNG stream position."""
    import ctypes as C
    import numpy as np
    from rfi_toolbox_b200 import _native
    from rfi_toolbox_b200.preprocessing.preprocessor import _keep_in_canonical_order
    C_, T_ = shape
    W = 3
    plan = _native.RfiPlan(dtype=0, magnitude=0, n_waterfalls=W, channels=C_, times=T_, patch=patch,
                           rotations=rot, stretch=1, norm_before=1, norm_after=0, flag_mode=1, sigma=5.0)
    n_groups = int(native_lib.rfi_plan_num_tiles(C.byref(plan)))
    n0 = int(native_lib.rfi_plan_num_patches(C.byref(plan)))
    skip = C_ <= patch and T_ <= patch
    padded = not skip and (C_ % patch or T_ % patch)
    nh, nw = (1, 1) if skip else (-(-C_ // patch), -(-T_ // patch))
    rng = np.random.default_rng(3)
    for frac, num_patches in [(0.7, None), (0.0, None), (1.0, 5), (0.3, 10_000)]:
        stats = np.zeros((n_groups, 22), dtype=np.int32)
        stats[:, 16] = (rng.random(n_groups) < frac) * rng.integers(1, 9, n_groups)
        nflag = stats[:, 16]
        keep = (nflag > 0) if padded else _keep_in_canonical_order((nflag > 0).reshape(W, nh, nw), rot)
        assert keep.size == n0
        np.random.seed(11)
        kept = np.flatnonzero(keep) if keep.any() else np.arange(n0)
        want = kept[np.random.permutation(len(kept))]
        if num_patches and num_patches < len(want):
            want = want[:num_patches]
        after = np.random.random(4)
        order, dest = np.empty(n0, dtype=np.int64), np.empty(n0, dtype=np.int64)
        np.random.seed(11)
        n_out = _native.plan_slots(plan, nflag, True, num_patches, order, dest)
        assert np.array_equal(np.random.random(4), after)
        assert n_out == len(want) and np.array_equal(order[:n_out], want)
        ref = np.full(n0, -1, dtype=np.int64)
        ref[want] = np.arange(len(want))
        assert np.array_equal(dest, ref)
    # inference mode: canonical order, RNG untouched
    np.random.seed(5)
    a = np.random.get_state()[2]
    n_out = _native.plan_slots(plan, None, False, None, order, dest)
    assert n_out == n0 and np.array_equal(order, np.arange(n0)) and np.array_equal(dest, np.arange(n0))
    assert np.random.get_state()[2] == a


def test_batch_writer_files_match_reference_format(tmp_path):
    """BatchWriter (batched_dataset.py:79-184): chunking into batch_###.pt + metadata.json."""
    import json
    import torch
    from rfi_toolbox_b200.datasets import BatchWriter, TorchDataset
    w = BatchWriter(tmp_path / "out", samples_per_batch=5)
    total = 0
    for n in (3, 4, 6):
        imgs = torch.arange(total, total + n, dtype=torch.float32)[:, None, None, None].expand(n, 8, 8, 3).contiguous()
        w.add_batch(TorchDataset(imgs, torch.ones((n, 8, 8), dtype=torch.uint8), {}))
        total += n
    w.finalize()
    files = sorted((tmp_path / "out").glob("batch_*.pt"))
    assert [f.name for f in files] == ["batch_000.pt", "batch_001.pt", "batch_002.pt", "batch_003.pt"]
    sizes = [len(torch.load(f)["images"]) for f in files]
    assert sizes == [5, 2, 5, 1]  # the reference flushes everything accumulated, in chunks of 5
    first = torch.cat([torch.load(f)["images"][:, 0, 0, 0] for f in files])
    assert torch.equal(first, torch.arange(13, dtype=torch.float32))
    meta = json.loads((tmp_path / "out" / "metadata.json").read_text())
    assert meta["num_samples"] == 13 and meta["num_batches"] == 4 and meta["samples_per_batch"] == 5
    assert meta["image_shape"] == [8, 8, 3] and meta["mask_shape"] == [8, 8]


def test_order_preserving_key_decode_matches_device_convention():
    """`evaluation.statistics._key_to_f32` (host side of the sharded radix select) inverts the device's
    order-preserving key map (csrc/rfi_common.cuh to_key): ascending float order == ascending key order."""
    from rfi_toolbox_b200.evaluation.statistics import _key_to_f32
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 1e3, 500), [0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 3.4e38, -3.4e38]]).astype(np.float32)
    bits = vals.view(np.uint32).astype(np.uint64)
    keys = np.where(bits >> 31 == 1, (~bits) & 0xffffffff, bits | 0x80000000)          # to_key
    order = np.argsort(keys, kind="stable")
    assert np.all(np.diff(vals[order].astype(np.float64)) >= 0)                          # keys sort like the values
    for k, v in zip(keys, vals):
        back = _key_to_f32(int(k))
        assert back == v and np.signbit(back) == np.signbit(v)
    assert np.isnan(_key_to_f32(0xffffffff))                                             # the excluded key decodes to NaN
