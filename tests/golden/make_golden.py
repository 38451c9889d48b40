#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the
build container.  The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and travel
to the GPU box, where /root/reference does not exist.

    CI=1 python tests/golden/make_golden.py

(`CI=1` makes the reference skip its multiprocessing pools, preprocessor.py:491-492.)

Each case stores: the sha256 of the input cube and mask (the arrays themselves are rebuilt
by tests.cubes.make_cube from the seed recorded in MANIFEST.json -- a checksum mismatch means
NumPy's generator stream changed and the fixtures must be regenerated here),
the reference's labels (bit-packed), the canonical order recovered from the labels is NOT
stored -- instead `perm_seed` lets the oracle replay the same np.random stream -- a sample of
8192 image values at fixed positions plus the float64 sum of every image channel, and the
reference's evaluate_segmentation / compute_ffi results on the same arrays.
"""
import hashlib
import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(ROOT))  # ahead of the reference: both trees have a `tests` package
os.environ.setdefault("CI", "1")

from tests.cubes import make_cube  # noqa: E402

CASES = {
    # name: (cube kwargs, use |z| as input, create_dataset kwargs, pass exact mask as flags)
    "real_sqrt_mad5": (dict(dtype=np.float32, seed=101), False,
                       dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False), False),
    "real_none_mad3_r2": (dict(dtype=np.float32, seed=102), False,
                          dict(patch_size=128, stretch=None, flag_sigma=3, use_custom_flags=False,
                               augmentation_rotations=2), False),
    "real_log10_mad5": (dict(dtype=np.float32, seed=103), False,
                        dict(patch_size=128, stretch="LOG10", flag_sigma=5, use_custom_flags=False), False),
    "real_sqrt_custom_after": (dict(dtype=np.float32, seed=104), False,
                               dict(patch_size=128, stretch="SQRT", use_custom_flags=True,
                                    normalize_after_stretch=True, num_patches=20), True),
    "complex_custom": (dict(dtype=np.complex64, seed=105), False,
                       dict(patch_size=128, use_custom_flags=True), True),
    "magnitude_sqrt_mad5": (dict(dtype=np.complex64, seed=106), True,
                            dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False), False),
    "real64_sqrt_mad5": (dict(dtype=np.float64, seed=107, n_bl=1), False,
                         dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False), False),
    "inference_special": (dict(dtype=np.float32, seed=108, special=True, n_bl=1), False,
                          dict(patch_size=128, stretch="SQRT", inference_mode=True), False),
    # BASELINE configs[4]'s literal call (big-tile path): |complex64|, P = 256, MAD sigma 3, no stretch
    "magnitude_p256_mad3": (dict(dtype=np.complex64, seed=109, channels=512, times=512), True,
                            dict(patch_size=256, stretch=None, flag_sigma=3, use_custom_flags=False), False),
    # dims that are not multiples of P: every ROTATED view is zero-padded bottom / right (:527-550)
    "padded_sqrt_mad5": (dict(dtype=np.float32, seed=110, channels=200, times=300), False,
                         dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False), False),
    # BASELINE configs[2]'s literal call: |complex64|, LOG10 stretch, MAD sigma 5 (exact-zero rows -> inf fill)
    "magnitude_log10_mad5": (dict(dtype=np.complex64, seed=111), True,
                             dict(patch_size=128, stretch="LOG10", flag_sigma=5, use_custom_flags=False), False),
}
PERM_SEED = 4242
N_SAMPLE = 8192


def main():
    import numpy
    import scipy
    import torch
    from rfi_toolbox.evaluation import compute_ffi, compute_statistics, evaluate_segmentation
    from rfi_toolbox.preprocessing import Preprocessor

    manifest = {"numpy": numpy.__version__, "scipy": scipy.__version__, "torch": torch.__version__,
                "reference": "rfi_toolbox 0.2.0 (/root/reference)", "perm_seed": PERM_SEED, "cases": {}}
    for name, (ck, use_abs, kw, with_flags) in CASES.items():
        ck = dict(ck)
        ck.setdefault("n_bl", 1)
        cube, mask = make_cube(**{"n_pol": 2, "channels": 256, "times": 256, **ck})
        data = np.abs(cube) if use_abs else cube
        np.random.seed(PERM_SEED)
        ds = Preprocessor(data, mask if with_flags else None).create_dataset(num_workers=0, **kw)
        images, labels = ds.images.numpy(), ds.labels.numpy()
        rng = np.random.default_rng(7)
        pos = rng.integers(0, images.size, size=min(N_SAMPLE, images.size))
        pred = labels.astype(bool)
        truth = pred ^ (np.random.default_rng(9).random(pred.shape) < 0.02)
        ev = evaluate_segmentation(pred, truth)
        ffi = compute_ffi(cube, mask)
        st = compute_statistics(cube, mask)
        out = dict(
            cube_sha256=np.array(hashlib.sha256(np.ascontiguousarray(cube).tobytes()).hexdigest()),
            mask_sha256=np.array(hashlib.sha256(np.packbits(mask).tobytes()).hexdigest()), use_abs=use_abs,
            labels=np.packbits(labels.astype(bool)), labels_shape=np.array(labels.shape),
            labels_max=np.int64(labels.max() if labels.size else 0),
            image_pos=pos, image_val=images.reshape(-1)[pos],
            image_channel_sum=np.nansum(images.astype(np.float64), axis=(0, 1, 2)),
            image_nan_count=np.int64(np.isnan(images).sum()),
            eval_keys=np.array(sorted(ev)), eval_vals=np.array([float(ev[k]) for k in sorted(ev)]),
            ffi_keys=np.array(sorted(ffi)), ffi_vals=np.array([ffi[k] for k in sorted(ffi)]),
            stat_keys=np.array(sorted(st)), stat_vals=np.array([float(st[k]) for k in sorted(st)]),
        )
        np.savez_compressed(HERE / f"{name}.npz", **out)
        manifest["cases"][name] = {"cube": ck | {"dtype": np.dtype(ck["dtype"]).name}, "abs": use_abs,
                                   "kwargs": kw, "flags": with_flags, "n_patches": int(len(labels)),
                                   "metadata": {k: (v if not isinstance(v, list) else [list(x) for x in v])
                                                for k, v in ds.metadata.items()}}
        print(name, images.shape, "flagged", int(labels.sum()))
    (HERE / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, default=str))


if __name__ == "__main__":
    main()
