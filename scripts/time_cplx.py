"""The reference's production call (synthetic_train_4k.yaml: 1024 x 1024 complex64 waterfalls, patch_size = 1024,
custom flags, 4 rotations): device-event time of create_dataset over a batch of samples."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from rfi_toolbox_b200 import Preprocessor
from rfi_toolbox_b200.utils.synth import device_cube
dev = torch.device("cuda", 0)
n_bl = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for patch in (1024, 256, 128):
    cube, mask = device_cube(n_bl, 4, 1024, 1024, seed=1234, device=dev)
    times = []
    for it in range(8):
        np.random.seed(0)
        pre = Preprocessor(cube, mask.bool())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        ds = pre.create_dataset(patch_size=patch, use_custom_flags=True)
        e1.record(); torch.cuda.synchronize()
        if it >= 2: times.append(e0.elapsed_time(e1))
        n = len(ds); del ds, pre
    ms = float(np.mean(times))
    print(f"complex branch, custom flags, P={patch}: {ms:.3f} ms per call, {cube.numel() / ms / 1e6:.1f} Gpix/s, {n} patches kept")
