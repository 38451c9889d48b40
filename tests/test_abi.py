"""The C-ABI library loads and exports every symbol include/rfi_b200.h declares (no compute
calls: this runs without a GPU), and the ctypes structs match the C layout."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "rfi_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfi_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(native_lib):
    from rfi_toolbox_b200 import _native
    names = _declared()
    assert len(names) >= 10
    raw = C.CDLL(str(_native.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in rfi_b200.h but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes binding"
    assert set(_native.SYMBOLS) == set(names)


def test_struct_layout_and_pure_host_calls(native_lib):
    from rfi_toolbox_b200 import _native
    assert native_lib.rfi_abi_version() == _native.ABI_VERSION
    assert C.sizeof(_native.RfiTileStat) == 88 and C.sizeof(_native.RfiStats) == 64 and C.sizeof(_native.RfiPlan) == 64
    plan = _native.RfiPlan(dtype=0, magnitude=0, n_waterfalls=8, channels=1024, times=2048, patch=128,
                           rotations=4, stretch=1, norm_before=1, norm_after=0, flag_mode=1, sigma=5.0)
    assert native_lib.rfi_plan_num_tiles(C.byref(plan)) == 8 * 8 * 16
    assert native_lib.rfi_plan_num_patches(C.byref(plan)) == 8 * 8 * 16 * 4
    assert native_lib.rfi_statistics_workspace_bytes() > 0
    # argument validation happens before any CUDA call
    plan.rotations = 3
    assert native_lib.rfi_tile_stats(C.byref(plan), None, None, None, None, None) == _native.RFI_E_INVALID
    assert b"rotations" in native_lib.rfi_last_error_string()
    # fast path (P = 128, dims divisible), float32: 64 KB of thread-private key scratch per tile; float64: none
    assert native_lib.rfi_plan_path(C.byref(plan)) == _native.RFI_PATH_FAST
    assert native_lib.rfi_plan_workspace_bytes(C.byref(plan)) == 8 * 8 * 16 * 65536
    plan.dtype = _native.RFI_F64
    assert native_lib.rfi_plan_workspace_bytes(C.byref(plan)) == 0
    plan.dtype = _native.RFI_F32
    # generic geometries: statistic groups, patches and workspace (host arithmetic only)
    plan.rotations, plan.patch = 4, 256
    assert native_lib.rfi_plan_path(C.byref(plan)) == _native.RFI_PATH_BIG
    assert native_lib.rfi_plan_num_tiles(C.byref(plan)) == 8 * 4 * 8
    assert native_lib.rfi_plan_num_patches(C.byref(plan)) == 8 * 4 * 8 * 4
    assert native_lib.rfi_plan_workspace_bytes(C.byref(plan)) > 0
    assert native_lib.rfi_tile_stats(C.byref(plan), None, None, None, None, None) == _native.RFI_E_INVALID
    assert b"NULL" in native_lib.rfi_last_error_string()
    plan.patch, plan.channels, plan.times = 100, 250, 330  # padded: one group per output patch
    assert native_lib.rfi_plan_num_patches(C.byref(plan)) == 8 * 4 * 3 * 4
    assert native_lib.rfi_plan_num_tiles(C.byref(plan)) == 8 * 4 * 3 * 4
    plan.patch = 512  # patchify skipped: one patch per rotated waterfall; R = 4 needs a square one
    assert native_lib.rfi_plan_num_patches(C.byref(plan)) == -1
    plan.rotations = 2
    assert native_lib.rfi_plan_num_patches(C.byref(plan)) == 8 * 2
    assert native_lib.rfi_plan_num_tiles(C.byref(plan)) == 8


def test_sass_is_sm100a(native_lib):
    """The shipped library carries sm_100a code only."""
    import shutil
    import subprocess
    from rfi_toolbox_b200 import _native
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
