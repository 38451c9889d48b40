#!/usr/bin/env python
"""Distance of the image channels from the float64 chain (oracle.images_exact64), for the CUDA path
and for the NumPy oracle, on the parity-test cubes.  Prints one line per case and channel:
    case  channel  max|numpy - exact|  max|gpu - exact|  max|gpu - numpy|
Run on the GPU box (optionally with RFI_B200_LIB pointing at another build of the library)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402  (checker)
from tests.cubes import make_cube  # noqa: E402

CASES = [
    ("f32 SQRT mad5", np.float32, False, dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)),
    ("f32 None mad5", np.float32, False, dict(stretch=None, flag_sigma=5, use_custom_flags=False)),
    ("f32 SQRT after", np.float32, False, dict(stretch="SQRT", flag_sigma=4, use_custom_flags=False, normalize_after_stretch=True)),
    ("f32 LOG10 mad5", np.float32, False, dict(stretch="LOG10", flag_sigma=5, use_custom_flags=False)),
    ("c64 mag SQRT", np.complex64, True, dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)),
    ("c64 mag LOG10", np.complex64, True, dict(stretch="LOG10", flag_sigma=5, use_custom_flags=False)),
    ("c64 mag SQRT P256", np.complex64, True, dict(patch_size=256, stretch="SQRT", flag_sigma=3, use_custom_flags=False)),
    ("c64 complex custom", np.complex64, False, dict(use_custom_flags=True)),
    ("f64 SQRT mad5", np.float64, False, dict(stretch="SQRT", flag_sigma=5, use_custom_flags=False)),
]


def main():
    import torch
    from rfi_toolbox_b200 import Preprocessor
    print(f"{'case':22s} ch  numpy-exact   gpu-exact     gpu-numpy    labels!=")
    for name, dtype, mag, kw in CASES:
        big = kw.get("patch_size", 128) == 256
        data, mask = make_cube(dtype=dtype, seed=3, channels=512 if big else 256, times=512 if big else 384)
        flags = mask if kw.get("use_custom_flags") else None
        np.random.seed(1)
        pre = Preprocessor(data, flags, magnitude=mag)
        ds = pre.create_dataset(**kw)
        torch.cuda.synchronize()
        np.random.seed(1)
        ods, inter = oracle.create_dataset(np.abs(data) if mag else data, flags, return_intermediates=True, **kw)
        assert np.array_equal(pre.order, inter["order"])
        exact = oracle.images_exact64(inter["processed"][inter["order"]])
        gpu = ds.images.cpu().numpy().astype(np.float64)
        ref = ods.images.astype(np.float64)
        nl = int((ds.labels.cpu().numpy() != ods.labels).sum())
        for c in range(3):
            e_np = np.nanmax(np.abs(ref[..., c] - exact[..., c]))
            e_gpu = np.nanmax(np.abs(gpu[..., c] - exact[..., c]))
            e_d = np.nanmax(np.abs(gpu[..., c] - ref[..., c]))
            print(f"{name:22s} {c}   {e_np:.3e}    {e_gpu:.3e}    {e_d:.3e}   {nl if c == 0 else ''}")


if __name__ == "__main__":
    main()
