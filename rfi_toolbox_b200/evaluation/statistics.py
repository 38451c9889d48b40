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


def compute_calcquality(data, flags, reference_data=None):
    """statistics.py:100-193 (lower is better): sensitivity, mean shift, std shift and
    over-flagging penalty from the statistics before / after flagging; the reductions
    (moments, max) run on the GPU, the formulas are the reference's Python-float ones."""
    if reference_data is not None:
        ref_st, _ = _run(reference_data, None)
        ref_stats = compute_statistics(reference_data, flags=None)
    else:
        ref_st, _ = _run(data, None)
        ref_stats = compute_statistics(data, flags=None)
    flag_stats = compute_statistics(data, flags=flags)
    rmean, rstd = ref_stats["mean"], ref_stats["std"]
    fmean, fstd = flag_stats["mean"], flag_stats["std"]
    pflag = flag_stats["flagged_fraction"] * 100
    if np.isnan(fmean) or np.isnan(fstd) or rstd < 1e-10:
        return {"calcquality": np.inf, "sensitivity": np.inf, "mean_shift": np.inf, "std_shift": np.inf,
                "overflagging_penalty": np.inf, "flagged_pct": float(pflag), "components": {}}
    rmax = ref_st.max
    maxdev = (rmax - rmean) / rstd
    fdiff = fmean - rmean
    sdiff = fstd - rstd
    a = abs(abs(maxdev) - 3)
    b = abs(fdiff) / rstd - 1
    c = abs(sdiff) / rstd
    d = max(0, (pflag - 70) / 10)
    calcquality = np.sqrt(a**2 + b**2 + c**2 + d**2)
    return {
        "calcquality": float(calcquality), "sensitivity": float(a), "mean_shift": float(b),
        "std_shift": float(c), "overflagging_penalty": float(d), "flagged_pct": float(pflag),
        "components": {"rmean": float(rmean), "rstd": float(rstd), "fmean": float(fmean), "fstd": float(fstd),
                       "rmax": float(rmax), "maxdev": float(maxdev), "fdiff": float(fdiff), "sdiff": float(sdiff)},
    }


def print_statistics_comparison(data, flags):
    """statistics.py:196-229 -- same text, same number formats."""
    b = compute_statistics(data, flags=None)
    a = compute_statistics(data, flags=flags)
    f = compute_ffi(data, flags)
    print("\n" + "=" * 60)
    print("Statistics Comparison (Before/After Flagging)")
    print("=" * 60)
    print("\nBefore Flagging:")
    print(f"  Mean:   {b['mean']:.4e}")
    print(f"  Median: {b['median']:.4e}")
    print(f"  Std:    {b['std']:.4e}")
    print(f"  MAD:    {b['mad']:.4e}")
    print(f"  Count:  {b['count']}")
    print(f"\nAfter Flagging ({a['flagged_fraction']*100:.2f}% flagged):")
    print(f"  Mean:   {a['mean']:.4e}")
    print(f"  Median: {a['median']:.4e}")
    print(f"  Std:    {a['std']:.4e}")
    print(f"  MAD:    {a['mad']:.4e}")
    print(f"  Count:  {a['count']}")
    print("\nFlagging Fidelity Index (FFI):")
    print(f"  FFI:            {f['ffi']:.4f}")
    print(f"  MAD Reduction:  {f['mad_reduction']:.4f}")
    print(f"  STD Reduction:  {f['std_reduction']:.4f}")


# ---------------------------------------------------------------------------------------------
# per-pair sweeps (extension; BASELINE config 4): one launch for a stack of patch pairs
_STATS_DTYPE = np.dtype([("mean", "f8"), ("median", "f8"), ("std", "f8"), ("mad", "f8"),
                         ("count", "i8"), ("n_flagged", "i8"), ("n_nan", "i8"), ("max", "f8")])


def _run_batch(data, flags):
    lib = _native.load()
    device = _device_of(data, flags)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported data dtype {d.dtype} (float32/64, complex64/128)")
    if d.ndim < 2:
        raise ValueError("batched statistics need a leading pair axis: data.shape = (N, ...)")
    n_seg = d.shape[0]
    seg = d.numel() // n_seg if n_seg else 0
    f = None
    if flags is not None:
        f = _flags_tensor(flags, device)
        if tuple(f.shape) != tuple(d.shape):
            raise IndexError("flags and data must have the same shape")
    if n_seg == 0:
        return np.zeros((0, 2), dtype=_STATS_DTYPE), seg
    with torch.cuda.device(device):
        out = torch.empty((n_seg, 2, _STATS_DTYPE.itemsize), dtype=torch.uint8, device=device)
        rc = lib.rfi_statistics_segmented(d.data_ptr(), _DTYPE_CODE[d.dtype], f.data_ptr() if f is not None else None,
                                          n_seg, seg, out.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_statistics_segmented")
        host = out.cpu().numpy()
    return host.view(_STATS_DTYPE).reshape(n_seg, 2), seg


def compute_statistics_batch(data, flags=None):
    """`compute_statistics(data[i], flags[i])` for every i in one launch (one CTA per pair,
    pairs of at most 16384 samples).  Returns a dict of NumPy arrays of length N."""
    st, seg = _run_batch(data, flags)
    col = st[:, 1 if flags is not None else 0]
    count = col["count"].astype(np.int64)
    empty = count == 0
    with np.errstate(invalid="ignore", divide="ignore"):
        frac = (col["n_flagged"] / float(seg)) if flags is not None else np.zeros(len(col))
    out = {k: np.where(empty, np.nan, col[k]) for k in ("mean", "median", "std", "mad")}
    out["count"] = count
    out["flagged_fraction"] = np.where(empty, 1.0, frac)
    return out


def compute_ffi_batch(data, flags, errors="raise"):
    """`compute_ffi(data[i], flags[i])` for every i (statistics.py:59-97) in one launch.

    Returns {'ffi','mad_reduction','std_reduction','flagged_fraction'} as float64 arrays whose
    entries equal the reference's per-pair results.  A pair with constant data (MAD or std of
    0 before flagging) makes the reference raise ZeroDivisionError; so does this function,
    naming the first such pair, unless `errors="nan"` (those entries become NaN)."""
    st, seg = _run_batch(data, flags)
    before, after = st[:, 0], st[:, 1]
    n = len(st)
    # statistics.py:77-78 -- all flagged (count 0) or NaN in the unflagged data
    guard = (after["count"] == 0) | np.isnan(after["mad"]) | np.isnan(after["std"])
    zero = ~guard & ((before["mad"] == 0) | (before["std"] == 0))
    if zero.any() and errors == "raise":
        raise ZeroDivisionError(f"float division by zero (pair {int(np.flatnonzero(zero)[0])}: constant data)")
    with np.errstate(invalid="ignore", divide="ignore"):
        mad_red = 1.0 - (after["mad"] / before["mad"])
        std_red = 1.0 - (after["std"] / before["std"])
        penalty = after["n_flagged"] / float(seg) if seg else np.zeros(n)
        ffi = (0.5 * mad_red + 0.5 * std_red) * (1.0 - 0.5 * penalty)
    bad = zero
    res = {"ffi": np.where(guard, 0.0, np.where(bad, np.nan, ffi)),
           "mad_reduction": np.where(guard, 0.0, np.where(bad, np.nan, mad_red)),
           "std_reduction": np.where(guard, 0.0, np.where(bad, np.nan, std_red)),
           "flagged_fraction": np.where(guard, 1.0, penalty)}
    return res


_PAIR_DTYPE = np.dtype([("ffi", "f8"), ("mad_reduction", "f8"), ("std_reduction", "f8"), ("flagged_fraction", "f8"),
                        ("tp", "u4"), ("fp", "u4"), ("fn", "u4"), ("status", "i4")])
_PINNED_RESULTS = {}


def _mask_u8(mask, device, what):
    """bool / uint8 masks as they are; anything else through `!= 0` (astype(bool), metrics.py:36)."""
    t = as_device_tensor(mask, device)
    if t.dtype == torch.bool:
        return t.contiguous().view(torch.uint8)
    if t.dtype == torch.uint8:
        return t.contiguous()
    return (t != 0).contiguous().view(torch.uint8)


def evaluate_pairs(data, pred, true, errors="raise"):
    """BASELINE config 4 as ONE launch: for every pair i of a stack (N, ...) the results of
    `compute_ffi(data[i], pred[i])` (statistics.py:59-97) and of `evaluate_segmentation(pred[i], true[i])`
    (metrics.py:155-172), with data, pred and true each read once (`rfi_pair_sweep`).

    data: float32 / complex64 (N, ...), at most 16384 samples per pair (other dtypes / sizes: use
    `compute_ffi_batch` + `evaluate_segmentation_batch`).  `pred` must be boolean, as the reference's
    `data[~flags]` demands.  Returns float64 arrays 'ffi', 'mad_reduction', 'std_reduction',
    'flagged_fraction', 'iou', 'precision', 'recall', 'f1', 'dice' and int64 'tp', 'fp', 'fn'.
    Constant data makes the reference raise ZeroDivisionError; so does this function, naming the first
    such pair, unless `errors="nan"`."""
    from .metrics import _ratio_arrays
    lib = _native.load()
    device = _device_of(data, pred, true)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype not in (torch.float32, torch.complex64):
        raise TypeError(f"evaluate_pairs: float32 / complex64 data (got {d.dtype})")
    if d.ndim < 2:
        raise ValueError("evaluate_pairs needs a leading pair axis: data.shape = (N, ...)")
    n = d.shape[0]
    seg = d.numel() // n if n else 0
    f = _flags_tensor(pred, device)
    t = _mask_u8(true, device, "true")
    if tuple(f.shape) != tuple(d.shape) or t.numel() != d.numel():
        raise IndexError("data, pred and true must have the same shape")
    if n == 0 or seg == 0:
        z = np.zeros(0)
        return {k: z for k in ("ffi", "mad_reduction", "std_reduction", "flagged_fraction", "iou", "precision",
                               "recall", "f1", "dice", "tp", "fp", "fn")}
    with torch.cuda.device(device):
        res = torch.empty((n, _PAIR_DTYPE.itemsize), dtype=torch.uint8, device=device)
        rc = lib.rfi_pair_sweep(d.contiguous().data_ptr(), _DTYPE_CODE[d.dtype], f.contiguous().data_ptr(), t.data_ptr(),
                                n, seg, None, res.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_pair_sweep")
        # results come down through a pinned buffer kept per device (48 B per pair)
        key = device.index
        host = _PINNED_RESULTS.get(key)
        if host is None or host.shape[0] < n:
            host = _PINNED_RESULTS[key] = torch.empty((n, _PAIR_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
        host[:n].copy_(res, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        r = host[:n].numpy().view(_PAIR_DTYPE).reshape(n).copy()
    zero = r["status"] == 2
    if zero.any() and errors == "raise":
        raise ZeroDivisionError(f"float division by zero (pair {int(np.flatnonzero(zero)[0])}: constant data)")
    out = {k: r[k].astype(np.float64) for k in ("ffi", "mad_reduction", "std_reduction", "flagged_fraction")}
    tp, fp, fn = (r[k].astype(np.int64) for k in ("tp", "fp", "fn"))
    out.update(_ratio_arrays(tp, fp, fn))
    return out
