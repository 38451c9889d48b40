"""GPUPreprocessor.create_raw_patches: the oracle port against the live reference on
the branch the reference can run, and the CUDA path against the oracle."""
import numpy as np
import pytest

import oracle
from tests.conftest import REFERENCE
from tests.cubes import make_cube


@pytest.mark.reference
def test_oracle_matches_reference_on_whole_waterfalls():
    """Upstream only runs when the waterfall is no larger than the patch (:885-890)."""
    import sys
    sys.path.insert(0, str(REFERENCE))
    from rfi_toolbox.preprocessing import GPUPreprocessor as Ref
    data, mask = make_cube(n_bl=3, n_pol=2, channels=64, times=96, dtype=np.complex64, seed=3)
    mask[1] = False  # blank waterfalls are dropped
    for flags in (mask, None):
        for kw in (dict(patch_size=128), dict(patch_size=128, num_patches=3), dict(patch_size=128, remove_blank=False)):
            np.random.seed(4)
            rp, rm = Ref(data, flags).create_raw_patches(num_workers=0, **kw)
            np.random.seed(4)
            op, om = oracle.create_raw_patches(data, flags, **kw)
            assert len(rp) == len(op)
            assert all(np.array_equal(a, b) for a, b in zip(rp, op))
            assert all(np.array_equal(a, b) for a, b in zip(rm, om))
    with pytest.raises(ValueError):
        oracle.create_raw_patches(np.abs(data))
    with pytest.raises(ValueError):
        Ref(np.abs(data))


@pytest.mark.reference
@pytest.mark.parametrize("shape,p", [((256, 384), 100), ((256, 384), 300), ((64, 96), 64), ((128, 128), 64)])
def test_oracle_padding_is_the_reference_worker_route(shape, p):
    """The tiling of the default route (num_workers > 0) is the reference's own `_patchify_single_waterfall`
    (preprocessor.py:46-112), which runs stand-alone: zero pad bottom / right, then P x P tiles in row-major order."""
    import sys
    sys.path.insert(0, str(REFERENCE))
    from rfi_toolbox.preprocessing.preprocessor import _patchify_single_waterfall
    data, mask = make_cube(n_bl=1, n_pol=2, channels=shape[0], times=shape[1], dtype=np.complex64, seed=8)
    np.random.seed(1)
    op, om = oracle.create_raw_patches(data, mask, patch_size=p, remove_blank=False, num_workers=4)
    order = np.random.RandomState(1).permutation(len(op))       # the one permutation the call drew
    ref_p, ref_m = [], []
    for w, m in zip(data[0], mask[0]):
        ref_p += _patchify_single_waterfall(w, p)[0]
        ref_m += _patchify_single_waterfall(m, p)[0]
    assert len(ref_p) == len(op)
    for k, src in enumerate(order):
        assert np.array_equal(op[k], ref_p[src]) and np.array_equal(om[k], ref_m[src])


CASES = [
    dict(patch_size=128), dict(patch_size=64, num_patches=7), dict(patch_size=128, remove_blank=False),
    dict(patch_size=100),                   # zero-padded to multiples of the patch size (the default route)
    dict(patch_size=100, num_workers=0),    # remainders dropped (patchify without padding)
    dict(patch_size=300),                   # one dimension below the patch size: padded up to it
    dict(patch_size=300, num_workers=0, remove_blank=False),   # ... or no tile at all
    dict(patch_size=512),   # whole waterfalls
]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("with_flags", [True, False])
def test_raw_patches_match_oracle(native_lib, kw, dtype, with_flags):
    import torch
    from rfi_toolbox_b200 import GPUPreprocessor
    data, mask = make_cube(n_bl=2, n_pol=2, channels=256, times=384, dtype=dtype, seed=5, special=not with_flags)
    if not with_flags:
        data[0, 1] = 0          # |z| > 0 is False everywhere -> blank tiles
    mask[1, 0] = False
    flags = mask if with_flags else None
    np.random.seed(9)
    op, om = oracle.create_raw_patches(data, flags, **kw)
    s_ref = np.random.get_state()[2]
    np.random.seed(9)
    pre = GPUPreprocessor(data, flags)
    gp, gm = pre.create_raw_patches(**kw)
    torch.cuda.synchronize()
    assert np.random.get_state()[2] == s_ref                    # same RNG consumption
    assert len(gp) == len(op) and gm.dtype == torch.bool
    gp, gm = gp.cpu().numpy(), gm.cpu().numpy()
    for k in range(len(op)):
        assert np.array_equal(gp[k], op[k], equal_nan=True) and np.array_equal(gm[k], om[k])


@pytest.mark.gpu
def test_raw_patches_errors(native_lib):
    from rfi_toolbox_b200 import GPUPreprocessor
    data, _ = make_cube(n_bl=1, n_pol=1, channels=64, times=64, dtype=np.complex64, seed=1)
    with pytest.raises(ValueError):
        GPUPreprocessor(np.abs(data))
    with pytest.raises(ValueError):
        GPUPreprocessor(data[0, 0])
    p, m = GPUPreprocessor(data[0]).create_raw_patches(patch_size=64)   # 3-D input gets a baseline axis
    assert p.shape == (1, 64, 64) and m.shape == (1, 64, 64)
