"""Routes of the phase-1 tiles (monotone / general, and why a tile left the monotone kernel) and the
two kernel times, for a bench workload:  python scripts/diag_stats.py [c2|c3|c5] [baselines]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import WORKLOADS  # noqa: E402
from rfi_toolbox_b200 import Preprocessor  # noqa: E402
from rfi_toolbox_b200.utils.synth import device_cube  # noqa: E402

w = dict(WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"])
if len(sys.argv) > 2:
    w["n_bl"] = int(sys.argv[2])
cube, mask = device_cube(w["n_bl"], w["n_pol"], w["channels"], w["times"], seed=1234, device="cuda")
kw = dict(patch_size=w["patch"], stretch=w["stretch"], flag_sigma=w["sigma"], use_custom_flags=False)
np.random.seed(0)
pre = Preprocessor(cube, None, magnitude=True)
pre.profile = True
for _ in range(3):
    ds = pre.create_dataset(**kw)
torch.cuda.synchronize()
st = pre.last_tile_stats.cpu().numpy().view(np.int32).reshape(-1, 22)
route = st[:, 17]
print(w["name"], "tiles", len(route), "route counts", {int(k): int(((route & 255) == k).sum()) for k in np.unique(route & 255)},
      "fallback reasons", {int(k): int(((route >> 8) == k).sum()) for k in np.unique(route >> 8)})
print("stats ms", pre.events["stats"][0].elapsed_time(pre.events["stats"][1]), "write ms",
      pre.events["write"][0].elapsed_time(pre.events["write"][1]))
