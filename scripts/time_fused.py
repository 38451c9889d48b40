"""Kernel A/B on one GPU through the C ABI: rfi_fused_patches (one launch) against rfi_tile_stats +
rfi_write_patches on the bench cube, with the destination slots of the all-kept case (identity), so that
both variants write every patch.  Device-event times."""
import ctypes as C, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from rfi_toolbox_b200 import _native
from rfi_toolbox_b200.utils.synth import device_cube

stretch = {"SQRT": 1, "None": 0, "LOG10": 2}[sys.argv[1] if len(sys.argv) > 1 else "SQRT"]
n_bl = int(sys.argv[2]) if len(sys.argv) > 2 else 45
lib = _native.load()
dev = torch.device("cuda", 0)
cube, _ = device_cube(n_bl, 4, 1024, 1024, seed=1234, device=dev)
plan = _native.RfiPlan(dtype=_native.RFI_C64, magnitude=1, n_waterfalls=n_bl * 4, channels=1024, times=1024, patch=128,
                       rotations=4, stretch=stretch, norm_before=1, norm_after=0, flag_mode=_native.RFI_FLAGS_MAD, sigma=5.0)
nt, n0 = int(lib.rfi_plan_num_tiles(C.byref(plan))), int(lib.rfi_plan_num_patches(C.byref(plan)))
ws = torch.empty(int(lib.rfi_plan_workspace_bytes(C.byref(plan))), dtype=torch.uint8, device=dev)
stats = torch.empty((nt, _native.TILE_STAT_BYTES), dtype=torch.uint8, device=dev)
dest = torch.arange(n0, dtype=torch.int64, device=dev)
images = torch.empty((n0, 128, 128, 3), dtype=torch.float32, device=dev)
labels = torch.empty((n0, 128, 128), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream

def fused():
    _native.check(lib.rfi_fused_patches(C.byref(plan), cube.data_ptr(), None, stats.data_ptr(), dest.data_ptr(),
                                        images.data_ptr(), labels.data_ptr(), st), "fused")
def two():
    _native.check(lib.rfi_tile_stats(C.byref(plan), cube.data_ptr(), None, stats.data_ptr(), ws.data_ptr(), st), "stats")
    _native.check(lib.rfi_write_patches(C.byref(plan), cube.data_ptr(), None, stats.data_ptr(), dest.data_ptr(),
                                        images.data_ptr(), labels.data_ptr(), ws.data_ptr(), st), "write")
res = {}
for name, fn in (("fused", fused), ("two", two), ("fused", fused), ("two", two)):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    chk = (float(images[::97].double().nan_to_num().sum()), int(labels[::97].sum()))
    print(name, round(ms, 3), "ms", "Gpix/s", round(cube.numel() / ms / 1e6, 1), "B/px60 frac", round(cube.numel() * 60 / ms / 1e6 / 6559.4, 3), chk)
