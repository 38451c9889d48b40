"""Load a golden fixture (tests/golden/*.npz) and rebuild its input arrays."""
import hashlib
import json
from pathlib import Path

import numpy as np

from tests.cubes import make_cube

GOLDEN = Path(__file__).resolve().parent / "golden"
MANIFEST = json.loads((GOLDEN / "MANIFEST.json").read_text())
CASE_NAMES = sorted(MANIFEST["cases"])


def load_case(name):
    info = MANIFEST["cases"][name]
    z = np.load(GOLDEN / f"{name}.npz")
    ck = dict(info["cube"])
    ck["dtype"] = np.dtype(ck["dtype"])
    cube, mask = make_cube(**{"n_pol": 2, "channels": 256, "times": 256, **ck})
    assert hashlib.sha256(np.ascontiguousarray(cube).tobytes()).hexdigest() == str(z["cube_sha256"]), \
        "make_cube no longer reproduces the fixture input (NumPy RNG stream changed?): regenerate tests/golden"
    assert hashlib.sha256(np.packbits(mask).tobytes()).hexdigest() == str(z["mask_sha256"])
    shape = tuple(int(v) for v in z["labels_shape"])
    labels = np.unpackbits(z["labels"])[: int(np.prod(shape))].reshape(shape).astype(np.uint8)
    return dict(
        info=info, cube=cube, mask=mask, data=np.abs(cube) if info["abs"] else cube,
        flags=mask if info["flags"] else None, kwargs=dict(info["kwargs"]), labels=labels,
        image_pos=z["image_pos"], image_val=z["image_val"], image_channel_sum=z["image_channel_sum"],
        image_nan_count=int(z["image_nan_count"]),
        evaluation=dict(zip(z["eval_keys"].tolist(), z["eval_vals"].tolist())),
        ffi=dict(zip(z["ffi_keys"].tolist(), z["ffi_vals"].tolist())),
        stats=dict(zip(z["stat_keys"].tolist(), z["stat_vals"].tolist())),
        perm_seed=MANIFEST["perm_seed"],
    )


# image values downstream of float32 log10 / complex abs / arctan2 are host-SIMD dependent
# (BASELINE.md section 6); labels of the LOG10-stretch case likewise.
HOST_DEPENDENT_LABELS = {"real_log10_mad5", "magnitude_log10_mad5"}
