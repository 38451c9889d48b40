"""Throughput of the other create_dataset variants on the bench cube (run on the GPU box)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from rfi_toolbox_b200 import Preprocessor
from rfi_toolbox_b200.utils.synth import device_cube
cube, mask = device_cube(45, 4, 1024, 1024, seed=1234, device='cuda')
mag = cube.abs()
def run(name, data, flags, magnitude, **kw):
    ts = []
    for it in range(8):
        np.random.seed(0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        pre = Preprocessor(data, flags, magnitude=magnitude); pre.profile = True
        ds = pre.create_dataset(**kw)
        e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ev = pre.events
            ts.append((e0.elapsed_time(e1), ev['stats'][0].elapsed_time(ev['stats'][1]), ev['write'][0].elapsed_time(ev['write'][1])))
        n = len(ds); del ds, pre
    t = np.mean(ts, axis=0)
    print(f"{name:58s} total {t[0]:6.3f} ms  stats {t[1]:6.3f}  write {t[2]:6.3f}  -> {data.numel()/t[0]/1e6:6.1f} Gpix/s  ({n} patches)")
run("complex64 magnitude, SQRT, MAD 5, R=4 (bench)", cube, None, True, patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
run("complex64 COMPLEX BRANCH, custom flags, R=4 (generator call)", cube, mask, False, patch_size=128, stretch=None, use_custom_flags=True, normalize_before_stretch=False)
run("float32 magnitudes, SQRT, MAD 5, R=4", mag, None, False, patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
run("float32 magnitudes, custom flags, no norm, R=4", mag, mask, False, patch_size=128, stretch=None, use_custom_flags=True, normalize_before_stretch=False)
run("float32 magnitudes, SQRT, MAD 5, R=1", mag, None, False, patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False, enable_augmentation=False)
run("complex64 magnitude, LOG10, MAD 5, R=4", cube, None, True, patch_size=128, stretch="LOG10", flag_sigma=5, use_custom_flags=False)
run("complex64 magnitude, SQRT, MAD 5, R=4, inference", cube, None, True, patch_size=128, stretch="SQRT", inference_mode=True)
# dims that are not multiples of P (zero padding after the rotation): R single-view plans over rotated, padded copies
cube2, _ = device_cube(45, 4, 1000, 1000, seed=1234, device='cuda')
run("complex64 magnitude, 1000 x 1000 (padded), SQRT, MAD 5, R=4", cube2, None, True, patch_size=128, stretch="SQRT", flag_sigma=5, use_custom_flags=False)
