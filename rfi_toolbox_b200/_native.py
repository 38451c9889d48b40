"""ctypes binding of librfi_b200.so (C ABI declared in include/rfi_b200.h).

There is no CPU fallback: if the library cannot be loaded, or no CUDA device is present,
every entry point raises.  `rfi_toolbox_b200.csrc.build.build()` (run by
`__graft_entry__.build()`) produces the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# RFI_B200_LIB: another build of the same library (kernel A/B experiments); default = the in-tree one
LIB_PATH = Path(os.environ.get("RFI_B200_LIB") or Path(__file__).resolve().parent / "_lib" / "librfi_b200.so")

RFI_F32, RFI_F64, RFI_C64, RFI_C128 = 0, 1, 2, 3
RFI_STRETCH_NONE, RFI_STRETCH_SQRT, RFI_STRETCH_LOG10 = 0, 1, 2
RFI_FLAGS_CUSTOM, RFI_FLAGS_MAD, RFI_FLAGS_INFERENCE = 0, 1, 2
RFI_E_INVALID, RFI_E_UNSUPPORTED, RFI_E_CUDA = -1, -2, -3
RFI_PATH_FAST, RFI_PATH_BIG, RFI_PATH_GENERIC = 0, 1, 2
ABI_VERSION = 6


class RfiPlan(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("magnitude", C.c_int32),
        ("n_waterfalls", C.c_int64), ("channels", C.c_int64), ("times", C.c_int64),
        ("patch", C.c_int32), ("rotations", C.c_int32), ("stretch", C.c_int32),
        ("norm_before", C.c_int32), ("norm_after", C.c_int32), ("flag_mode", C.c_int32),
        ("sigma", C.c_double),
    ]


class RfiTileStat(C.Structure):
    _fields_ = [
        ("median_before", C.c_double), ("inf_fill", C.c_double), ("median_after", C.c_double),
        ("centre", C.c_double), ("mad", C.c_double), ("thr_lo", C.c_double), ("thr_hi", C.c_double),
        ("n_valid", C.c_int32), ("n_inf", C.c_int32), ("n_flagged", C.c_int32), ("route", C.c_int32),
        ("raw_lo", C.c_double), ("raw_hi", C.c_double),
    ]


class RfiStats(C.Structure):
    _fields_ = [
        ("mean", C.c_double), ("median", C.c_double), ("std", C.c_double), ("mad", C.c_double),
        ("count", C.c_int64), ("n_flagged", C.c_int64), ("n_nan", C.c_int64), ("max", C.c_double),
    ]


class RfiPairResult(C.Structure):
    _fields_ = [
        ("ffi", C.c_double), ("mad_reduction", C.c_double), ("std_reduction", C.c_double),
        ("flagged_fraction", C.c_double), ("iou", C.c_double), ("precision", C.c_double), ("recall", C.c_double),
        ("f1", C.c_double), ("dice", C.c_double), ("tp", C.c_uint32), ("fp", C.c_uint32), ("fn", C.c_uint32),
        ("status", C.c_int32),
    ]


class RfiSynth(C.Structure):
    _fields_ = [
        ("channels", C.c_int64), ("times", C.c_int64), ("n_pol", C.c_int32), ("enable_bandpass", C.c_int32),
        ("bandpass_order", C.c_int32), ("n_bands", C.c_int32), ("n_sweeps", C.c_int32),
        ("noise_level", C.c_float), ("pol_corr", C.c_float), ("seed", C.c_uint64),
    ]


TILE_STAT_BYTES = C.sizeof(RfiTileStat)
assert TILE_STAT_BYTES == 88

# every symbol include/rfi_b200.h declares: name -> (restype, argtypes)
_VP, _I, _I64 = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = {
    "rfi_plan_path": (_I, [C.POINTER(RfiPlan)]),
    "rfi_plan_num_tiles": (_I64, [C.POINTER(RfiPlan)]),
    "rfi_plan_num_patches": (_I64, [C.POINTER(RfiPlan)]),
    "rfi_plan_workspace_bytes": (C.c_size_t, [C.POINTER(RfiPlan)]),
    "rfi_tile_stats": (_I, [C.POINTER(RfiPlan), _VP, _VP, _VP, _VP, _VP]),
    "rfi_write_patches": (_I, [C.POINTER(RfiPlan), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rfi_plan_fusable": (_I, [C.POINTER(RfiPlan)]),
    "rfi_fused_patches": (_I, [C.POINTER(RfiPlan), _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rfi_confusion_counts": (_I, [_VP, _I, _I, _VP, _I, _I, _I64, _VP, _VP]),
    "rfi_confusion_counts_allreduce": (_I, [_VP, _I, _I, _VP, _I, _I, _I64, _VP, _I, _I, C.c_uint64, _VP, _VP]),
    "rfi_peer_alloc": (_I, [C.POINTER(C.c_void_p), _VP]),
    "rfi_peer_open": (_I, [_VP, C.POINTER(C.c_void_p)]),
    "rfi_peer_close": (_I, [_VP]),
    "rfi_peer_free": (_I, [_VP]),
    "rfi_confusion_counts_segmented": (_I, [_VP, _I, _I, _VP, _I, _I, _I64, _I64, _VP, _VP]),
    "rfi_statistics_workspace_bytes": (C.c_size_t, []),
    "rfi_statistics": (_I, [_VP, _I, _VP, _I64, _VP, _VP, _VP]),
    "rfi_statistics2_workspace_bytes": (C.c_size_t, [_I, _I64]),
    "rfi_statistics2": (_I, [_VP, _I, _VP, _I64, _VP, _VP, _VP]),
    "rfi_statistics_shard_begin": (_I, [_VP, _I, _VP, _I64, _VP, _VP, _VP]),
    "rfi_statistics_shard_count": (_I, [_VP, _I64, _VP, _I, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "rfi_statistics_segmented": (_I, [_VP, _I, _VP, _I64, _I64, _VP, _VP]),
    "rfi_pair_sweep": (_I, [_VP, _I, _VP, _VP, _I64, _I64, _VP, _VP, _VP]),
    "rfi_legacy_permutation": (_I, [_VP, C.POINTER(C.c_int32), _I64, _VP]),
    "rfi_plan_slots": (_I, [C.POINTER(RfiPlan), _VP, _I64, _I, _VP, C.POINTER(C.c_int32), _I64, _VP, _VP,
                            C.POINTER(C.c_int64)]),
    "rfi_synth_waterfalls": (_I, [C.POINTER(RfiSynth), _I64, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rfi_raw_num_tiles": (_I64, [_I, _I64, _I64, _I64, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rfi_raw_tile_counts": (_I, [_VP, _I, _VP, _I64, _I64, _I64, C.c_int32, _VP, _VP]),
    "rfi_raw_gather": (_I, [_VP, _I, _VP, _I64, _I64, _I64, C.c_int32, _VP, _VP, _VP, _VP]),
    "rfi_rotate_pad": (_I, [_VP, _VP, _I, _I64, _I64, _I64, _I64, _I64, _I, _VP]),
    "rfi_processed_patches": (_I, [C.POINTER(RfiPlan), _VP, _VP, _VP, _I64, _VP, _VP]),
    "rfi_downcast": (_I, [_VP, _VP, _I, _I64, _VP]),
    "rfi_selftest_sqrt_unit": (_I, [_VP, _VP]),
    "rfi_selftest_cabs_fast": (_I, [_VP, _VP]),
    "rfi_last_error_string": (C.c_char_p, []),
    "rfi_abi_version": (_I, []),
}


class NativeError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the ctypes handle.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python -m rfi_toolbox_b200.csrc.build` "
            "(nvcc, sm_100a).  rfi_toolbox_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype, fn.argtypes = res, args
    if lib.rfi_abi_version() != ABI_VERSION:
        raise NativeError(f"ABI mismatch: library {lib.rfi_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def _global_mt19937():
    """(key pointer, pos pointer) INTO the global legacy generator's MT19937 state -- the
    native helpers then advance the very stream `np.random.permutation` would -- or None when
    the global generator is not MT19937 / does not expose its state (the callers then go
    through get_state / set_state)."""
    import numpy as np
    try:
        bg = np.random.mtrand._rand._bit_generator
        if type(bg).__name__ != "MT19937":
            return None
        addr = int(bg.ctypes.state_address)  # struct { uint32_t key[624]; int pos; }
        return bg, addr, addr + 624 * 4
    except Exception:
        return None


def legacy_permutation(n: int):
    """`np.random.permutation(n)` of the GLOBAL legacy generator (same values, same stream
    position afterwards), computed by the native helper.  Falls back to NumPy itself when the
    global generator is not MT19937."""
    import numpy as np

    if n < 2:
        return np.random.permutation(n)
    direct = _global_mt19937()
    if direct is not None:
        bg, kaddr, paddr = direct
        out = np.empty(n, dtype=np.int64)
        with bg.lock:
            rc = load().rfi_legacy_permutation(kaddr, C.cast(paddr, C.POINTER(C.c_int32)), n, out.ctypes.data)
        check(rc, "rfi_legacy_permutation")
        return out
    state = np.random.get_state()
    if state[0] != "MT19937":
        return np.random.permutation(n)
    key = np.array(state[1], dtype=np.uint32, copy=True)
    pos = C.c_int32(int(state[2]))
    out = np.empty(n, dtype=np.int64)
    rc = load().rfi_legacy_permutation(key.ctypes.data, C.byref(pos), n, out.ctypes.data)
    check(rc, "rfi_legacy_permutation")
    np.random.set_state(("MT19937", key, int(pos.value), state[3], state[4]))
    return out


def plan_slots(plan, n_flagged, shuffle, num_patches, order, dest):
    """`rfi_plan_slots` on NumPy buffers, drawing from / advancing the GLOBAL legacy generator
    exactly like `np.random.permutation(n_kept)`.  `n_flagged` int32 [groups] (any stride) or
    None; `order`, `dest` int64 [n_patches] outputs.  Returns n_out."""
    import numpy as np

    lib = load()
    n_out = C.c_int64(0)
    fptr, stride = (None, 0) if n_flagged is None else (n_flagged.ctypes.data, n_flagged.strides[0])
    direct = _global_mt19937() if shuffle else None
    if direct is not None:
        bg, kaddr, paddr = direct
        with bg.lock:
            rc = lib.rfi_plan_slots(C.byref(plan), fptr, stride, 1, kaddr, C.cast(paddr, C.POINTER(C.c_int32)),
                                    int(num_patches) if num_patches else 0, order.ctypes.data, dest.ctypes.data,
                                    C.byref(n_out))
        check(rc, "rfi_plan_slots")
        return int(n_out.value)
    state = np.random.get_state() if shuffle else None
    if shuffle and state[0] != "MT19937":
        raise NativeError("the global NumPy generator is not MT19937; cannot reproduce the reference shuffle")
    key = np.array(state[1], dtype=np.uint32, copy=True) if shuffle else None
    pos = C.c_int32(int(state[2]) if shuffle else 0)
    rc = lib.rfi_plan_slots(C.byref(plan), fptr, stride, int(bool(shuffle)),
                            key.ctypes.data if shuffle else None, C.byref(pos),
                            int(num_patches) if num_patches else 0, order.ctypes.data, dest.ctypes.data,
                            C.byref(n_out))
    check(rc, "rfi_plan_slots")
    if shuffle:
        np.random.set_state(("MT19937", key, int(pos.value), state[3], state[4]))
    return int(n_out.value)


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = load().rfi_last_error_string().decode("utf-8", "replace")
    if rc == RFI_E_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == RFI_E_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise NativeError(f"{what}: {msg}")
