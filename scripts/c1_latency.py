"""configs[0]: one 1024 x 1024 waterfall (4 pols) -- latency of create_dataset + evaluate_segmentation."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rfi_toolbox_b200 import Preprocessor, evaluate_segmentation
from rfi_toolbox_b200.utils.synth import device_cube
cube, mask = device_cube(1, 4, 1024, 1024, seed=1234, device='cuda')
for name, data, flags, mag, kw in [
    ("real branch (|z| fused), SQRT, MAD 5", cube, None, True, dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False)),
    ("literal: complex in, custom flags", cube, mask, False, dict(patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=True))]:
    ts = []
    for it in range(30):
        np.random.seed(0); torch.cuda.synchronize(); t0 = time.perf_counter()
        ds = Preprocessor(data, flags, magnitude=mag).create_dataset(**kw)
        m = evaluate_segmentation(ds.labels, ds.labels)
        ts.append(time.perf_counter() - t0)
    t = np.median(ts[5:])
    print(f"{name:40s} {1e3*t:.3f} ms per call  ({cube.numel()/t/1e9:.2f} Gpix/s, {len(ds)} patches)")
