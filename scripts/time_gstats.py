"""compute_ffi / compute_statistics over the 45-baseline bench cube (188.7 M complex64 samples): wall time
per call (device work + the 128-byte result download), against a device-to-device copy of the cube."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from rfi_toolbox_b200 import compute_ffi, compute_statistics
from rfi_toolbox_b200.utils.synth import device_cube
dev = torch.device("cuda", 0)
cube, mask = device_cube(45, 4, 1024, 1024, seed=1234, device=dev)
mask = mask.bool()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r
dst = torch.empty_like(cube)
ms_copy, _ = timed(lambda: dst.copy_(cube))
ms_ffi, ffi = timed(lambda: compute_ffi(cube, mask))
ms_st, st = timed(lambda: compute_statistics(cube, mask))
print(f"copy {ms_copy:.3f} ms   compute_ffi {ms_ffi:.3f} ms ({ms_ffi / ms_copy:.2f} x copy)   compute_statistics {ms_st:.3f} ms")
print(ffi, st)
