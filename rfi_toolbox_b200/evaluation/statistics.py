"""Flagging statistics -- drop-in for rfi_toolbox/evaluation/statistics.py:10-97.

`compute_statistics(data, flags=None)` and `compute_ffi(data, flags)` keep the reference's
signatures, dict keys, Python-float results and edge cases (all-flagged / NaN guard,
bool-only flags, ZeroDivisionError on constant data).  The |z| pass, the boolean gather,
the moments and the exact median / MAD selections run on the GPU (`rfi_statistics`,
csrc/rfi_stats.cu); the final FFI arithmetic is the reference's Python-float formula.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _native
from ..utils.device import as_device_tensor, current_stream_ptr, require_cuda

_DTYPE_CODE = {
    torch.float32: _native.RFI_F32, torch.float64: _native.RFI_F64,
    torch.complex64: _native.RFI_C64, torch.complex128: _native.RFI_C128,
}


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return require_cuda()


def _flags_tensor(flags, device):
    """The reference indexes with `~flags` (statistics.py:33): only boolean masks work."""
    if isinstance(flags, torch.Tensor):
        if flags.dtype != torch.bool:
            raise IndexError("flags must be a boolean mask (the reference evaluates data[~flags])")
    elif np.asarray(flags).dtype != np.bool_:
        raise IndexError("flags must be a boolean mask (the reference evaluates data[~flags])")
    return as_device_tensor(flags, device).view(torch.uint8)


def _run(data, flags):
    lib = _native.load()
    device = _device_of(data, flags)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported data dtype {d.dtype} (float32/64, complex64/128)")
    f = None
    if flags is not None:
        f = _flags_tensor(flags, device)
        if f.numel() != d.numel():
            raise IndexError("flags and data must have the same shape")
    with torch.cuda.device(device):
        ws = torch.empty(int(lib.rfi_statistics_workspace_bytes()), dtype=torch.uint8, device=device)
        out = torch.empty(C.sizeof(_native.RfiStats), dtype=torch.uint8, device=device)
        rc = lib.rfi_statistics(d.data_ptr(), _DTYPE_CODE[d.dtype], f.data_ptr() if f is not None else None,
                                d.numel(), out.data_ptr(), ws.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_statistics")
        host = out.cpu().numpy().tobytes()
    st = _native.RfiStats.from_buffer_copy(host)
    return st, d.numel()


def compute_mad(data):
    """statistics.py:10-13 (median absolute deviation, scale 1.0)."""
    st, _ = _run(data, None)
    return _scalar(st.mad, data)


def _scalar(v, data):
    return v


def compute_statistics(data, flags=None):
    """statistics.py:16-56."""
    st, n = _run(data, flags)
    if flags is not None:
        flagged_fraction = st.n_flagged / n if n else float("nan")
    else:
        flagged_fraction = 0.0
    if st.count == 0:
        return {"mean": np.nan, "median": np.nan, "std": np.nan, "mad": np.nan,
                "count": 0, "flagged_fraction": 1.0}
    return {
        "mean": float(st.mean), "median": float(st.median), "std": float(st.std),
        "mad": float(st.mad), "count": int(st.count), "flagged_fraction": float(flagged_fraction),
    }


def compute_ffi(data, flags):
    """statistics.py:59-97."""
    before = compute_statistics(data, flags=None)
    after = compute_statistics(data, flags=flags)
    if np.isnan(after["mad"]) or np.isnan(after["std"]):
        return {"ffi": 0.0, "mad_reduction": 0.0, "std_reduction": 0.0, "flagged_fraction": 1.0}
    mad_reduction = 1.0 - (after["mad"] / before["mad"])
    std_reduction = 1.0 - (after["std"] / before["std"])
    penalty = after["flagged_fraction"]
    ffi = (0.5 * mad_reduction + 0.5 * std_reduction) * (1.0 - 0.5 * penalty)
    return {"ffi": float(ffi), "mad_reduction": float(mad_reduction),
            "std_reduction": float(std_reduction), "flagged_fraction": float(penalty)}
