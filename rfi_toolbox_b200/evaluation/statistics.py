"""Flagging statistics -- drop-in for rfi_toolbox/evaluation/statistics.py:10-97.

`compute_statistics(data, flags=None)` and `compute_ffi(data, flags)` keep the reference's
signatures, dict keys, Python-float results and edge cases (all-flagged / NaN guard,
bool-only flags, ZeroDivisionError on constant data).  The |z| pass, the boolean gather,
the moments and the exact median / MAD selections run on the GPU (`rfi_statistics2`,
csrc/rfi_gstats.cu: three passes over the cube for both sets; float64 input: csrc/rfi_stats.cu);
the final FFI arithmetic is the reference's Python-float formula.

Where the values can differ from the reference's (tolerance class, tests/test_gpu_metrics.py):
  * median / MAD / max: bit-identical (exact order statistics in the data's precision);
  * mean / std: float64 accumulation rounded once to the data's precision, where NumPy sums
    pairwise IN float32 for float32 data -- equal to ~1e-7 of mean(|x|), not to the last bit;
  * `flags` must have `data`'s number of elements (the reference's `data[~flags]` also accepts
    lower-rank boolean masks); integer data is promoted to float64 as NumPy promotes it.
Extensions: `group=` (cube sharded over ranks by baseline), `*_batch` / `evaluate_pairs`
(per-pair sweeps in one launch).
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


def _match_flags(f, d):
    """`data[~flags]` (statistics.py:33) with NumPy's boolean-mask rule: the mask covers the LEADING dimensions of
    the data -- all of them, or fewer (a mask per baseline / per waterfall drops whole slices).  Returns a mask
    with one byte per sample."""
    if tuple(f.shape) == tuple(d.shape):
        return f
    if f.ndim < d.ndim and tuple(f.shape) == tuple(d.shape[:f.ndim]):
        return f.reshape(f.shape + (1,) * (d.ndim - f.ndim)).expand(d.shape).contiguous()
    raise IndexError(f"boolean index did not match indexed array: flags {tuple(f.shape)}, data {tuple(d.shape)}")


def _run(data, flags):
    lib = _native.load()
    device = _device_of(data, flags)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported data dtype {d.dtype} (float32/64, complex64/128)")
    f = None
    if flags is not None:
        f = _match_flags(_flags_tensor(flags, device), d)
    with torch.cuda.device(device):
        ws = torch.empty(int(lib.rfi_statistics_workspace_bytes()), dtype=torch.uint8, device=device)
        out = torch.empty(C.sizeof(_native.RfiStats), dtype=torch.uint8, device=device)
        rc = lib.rfi_statistics(d.data_ptr(), _DTYPE_CODE[d.dtype], f.data_ptr() if f is not None else None,
                                d.numel(), out.data_ptr(), ws.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_statistics")
        host = out.cpu().numpy().tobytes()
    st = _native.RfiStats.from_buffer_copy(host)
    return st, d.numel()


def _run_both(data, flags):
    """(stats of all samples, stats of the unflagged samples, n) from ONE call (`rfi_statistics2`: the cube is
    read once, two more passes over a 4 B / px scratch); flags None -> the second equals the first."""
    lib = _native.load()
    device = _device_of(data, flags)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
        d = d.to(torch.float64)   # np.mean / np.median / np.std of integer data are float64
    if d.dtype not in _DTYPE_CODE:
        raise TypeError(f"unsupported data dtype {d.dtype} (float32/64, complex64/128 or an integer type)")
    f = None
    if flags is not None:
        f = _match_flags(_flags_tensor(flags, device), d)
    code = _DTYPE_CODE[d.dtype]
    with torch.cuda.device(device):
        ws = torch.empty(int(lib.rfi_statistics2_workspace_bytes(code, d.numel())), dtype=torch.uint8, device=device)
        out = torch.empty(2 * C.sizeof(_native.RfiStats), dtype=torch.uint8, device=device)
        rc = lib.rfi_statistics2(d.data_ptr(), code, f.data_ptr() if f is not None else None, d.numel(),
                                 out.data_ptr(), ws.data_ptr(), current_stream_ptr(device))
        _native.check(rc, "rfi_statistics2")
        host = out.cpu().numpy().tobytes()
    n = C.sizeof(_native.RfiStats)
    before, after = _native.RfiStats.from_buffer_copy(host[:n]), _native.RfiStats.from_buffer_copy(host[n:])
    if before.count < 0:   # a sampled bracket missed its rank: the exact radix passes instead
        before, _ = _run(d, None)
        after = _run(d, flags)[0] if flags is not None else before
    return before, after, d.numel()


class _Stats:
    """Field-compatible with `_native.RfiStats` (host-side result of the sharded path)."""
    __slots__ = ("mean", "median", "std", "mad", "count", "n_flagged", "n_nan", "max")


def _key_to_f32(k):
    """Order-preserving key (csrc/rfi_common.cuh to_key / from_key) -> np.float32."""
    k = int(k) & 0xffffffff
    bits = (k ^ 0x80000000) if (k & 0x80000000) else (~k & 0xffffffff)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def _run_both_sharded(data, flags, group):
    """`_run_both` over the UNION of the ranks' shards (baseline-sharded cube): pass A per rank,
    moments summed over `group`, order statistics by an MSB-first radix select on the ranks' key scratches
    whose per-pass counts are all-reduced (`rfi_statistics_shard_begin / _count`).  float32 / complex64."""
    import torch.distributed as dist
    lib = _native.load()
    device = _device_of(data, flags)
    require_cuda(device)
    d = as_device_tensor(data, device)
    if d.dtype not in (torch.float32, torch.complex64):
        raise TypeError(f"sharded statistics: float32 / complex64 data (got {d.dtype})")
    f = None
    if flags is not None:
        f = _match_flags(_flags_tensor(flags, device), d)
    g = None if group is True else group
    code, n_loc = _DTYPE_CODE[d.dtype], d.numel()
    fptr = f.data_ptr() if f is not None else None
    with torch.cuda.device(device):
        stream = current_stream_ptr(device)
        ws = torch.empty(int(lib.rfi_statistics2_workspace_bytes(code, n_loc)), dtype=torch.uint8, device=device)
        state = torch.zeros(8, dtype=torch.float64, device=device)
        _native.check(lib.rfi_statistics_shard_begin(d.data_ptr() if n_loc else None, code, fptr, n_loc, ws.data_ptr(),
                                                     state.data_ptr(), stream), "rfi_statistics_shard_begin")
        dist.all_reduce(state[:6], op=dist.ReduceOp.SUM, group=g)
        dist.all_reduce(state[6:], op=dist.ReduceOp.MAX, group=g)
        st = state.cpu().numpy()
        n_all, n_flag = int(st[0]), int(st[1])
        n = [n_all, n_all - n_flag]
        n_nan = [int(st[2]), int(st[3])]
        mean = [np.float32(st[4 + s] / n[s]) if n[s] else np.float32(0) for s in range(2)]
        n_valid = [n[s] - n_nan[s] for s in range(2)]
        k1 = [(n_valid[s] - 1) >> 1 if n_valid[s] else 0 for s in range(2)]
        k2 = [n_valid[s] >> 1 for s in range(2)]
        counts = torch.zeros(32, dtype=torch.int64, device=device)
        sumsq = torch.zeros(2, dtype=torch.float64, device=device)
        F2, U2 = C.c_float * 2, C.c_uint32 * 2

        def select(dev_mode, centre, want_moments):
            prefix = [0, 0]
            for p, shift in enumerate(range(28, -1, -4)):
                _native.check(lib.rfi_statistics_shard_count(
                    fptr, n_loc, ws.data_ptr(), 0, dev_mode, F2(*centre), F2(*mean), U2(*prefix), shift, counts.data_ptr(),
                    sumsq.data_ptr() if (want_moments and p == 0) else None, stream), "rfi_statistics_shard_count")
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=g)
                c = counts.cpu().numpy()
                for s in range(2):
                    prefix[s] |= int(np.count_nonzero(c[s * 16:s * 16 + 15] <= k1[s])) << shift
            _native.check(lib.rfi_statistics_shard_count(fptr, n_loc, ws.data_ptr(), 1, dev_mode, F2(*centre), F2(*mean),
                                                         U2(*prefix), 0, counts.data_ptr(), None, stream),
                          "rfi_statistics_shard_count")
            cle, nxt = counts[[0, 16]].clone(), counts[[1, 17]].clone()
            dist.all_reduce(cle, op=dist.ReduceOp.SUM, group=g)
            dist.all_reduce(nxt, op=dist.ReduceOp.MIN, group=g)   # int64 view of the uint64 keys: keys < 2^32
            cle, nxt = cle.cpu().numpy(), nxt.cpu().numpy()
            out = []
            for s in range(2):
                if n_valid[s] == 0:
                    out.append(np.float32(np.nan))
                    continue
                a = _key_to_f32(prefix[s])
                b = a if (k2[s] == k1[s] or k2[s] < int(cle[s])) else _key_to_f32(int(nxt[s]))
                out.append(a if (n_valid[s] & 1) else np.float32(np.float32(a + b) * np.float32(0.5)))
            return out

        with np.errstate(invalid="ignore", over="ignore"):
            med = select(0, [0.0, 0.0], True)
            dist.all_reduce(sumsq, op=dist.ReduceOp.SUM, group=g)
            ssq = sumsq.cpu().numpy()
            centre = [float(m) if np.isfinite(m) else 0.0 for m in med]
            mad = select(1, centre, False)
    res = []
    for s in range(2):
        o = _Stats()
        o.count, o.n_flagged, o.n_nan = n[s], (n_flag if s == 1 else 0), n_nan[s]
        o.mean = o.median = o.std = o.mad = o.max = float("nan")
        if n[s]:
            o.mean = float(mean[s])
            o.std = float(np.sqrt(np.float32(ssq[s] / n[s])))
            o.max = float(_key_to_f32(int(st[6 + s])))
            if n_nan[s] == 0:
                o.median = float(med[s])
                o.mad = float(mad[s]) if np.isfinite(med[s]) else float("nan")
        res.append(o)
    return res[0], res[1], n_all


def compute_mad(data):
    """statistics.py:10-13 (median absolute deviation, scale 1.0)."""
    st, _, _ = _run_both(data, None)
    return _scalar(st.mad, data)


def _scalar(v, data):
    return v


def compute_statistics(data, flags=None, group=None):
    """statistics.py:16-56.  `group` (extension; `True` = the default process group): `data` / `flags` are this
    rank's baseline shard and the statistics are those of the whole cube -- every rank gets the same dict."""
    before, after, n = _run_both(data, flags) if group is None else _run_both_sharded(data, flags, group)
    out = _stats_dict(after if flags is not None else before, n, flags is not None)
    if flags is not None and flags.ndim < data.ndim and out["count"]:
        # a mask over the leading dimensions: `count` is len(data[~flags]) (:52), the number of SLICES kept
        out["count"] //= int(np.prod(tuple(data.shape)[flags.ndim:]))
    return out


def _stats_dict(st, n, flagged):
    if flagged:
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


def compute_ffi(data, flags, group=None):
    """statistics.py:59-97 (both compute_statistics calls of :73-74 from one pass over the cube).  `group`: as
    for `compute_statistics` -- the FFI of a cube sharded over the ranks by baseline."""
    b, a, n = _run_both(data, flags) if group is None else _run_both_sharded(data, flags, group)
    before, after = _stats_dict(b, n, False), _stats_dict(a, n, flags is not None)
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
    b, a, n = _run_both(data, flags)
    flag_stats = _stats_dict(a, n, flags is not None)
    if reference_data is not None:
        ref_st, _, nr = _run_both(reference_data, None)
        ref_stats = _stats_dict(ref_st, nr, False)
    else:
        ref_st, ref_stats = b, _stats_dict(b, n, False)
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
    sb, sa, n = _run_both(data, flags)
    b, a = _stats_dict(sb, n, False), _stats_dict(sa, n, flags is not None)
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
    if seg > 16384:
        # pairs larger than one CTA's shared memory: the whole-cube routine per pair (three passes each)
        host = np.zeros((n_seg, 2), dtype=_STATS_DTYPE)
        dd, ff = d.reshape(n_seg, -1), (f.reshape(n_seg, -1).view(torch.bool) if f is not None else None)
        for i in range(n_seg):
            b, a, _ = _run_both(dd[i], ff[i] if ff is not None else None)
            for col, st in ((0, b), (1, a)):
                for k in _STATS_DTYPE.names:
                    host[i, col][k] = getattr(st, k)
            host[i, 0]["n_flagged"] = 0
        return host, seg
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
                        ("iou", "f8"), ("precision", "f8"), ("recall", "f8"), ("f1", "f8"), ("dice", "f8"),
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
        # 88 B per pair = 11 x 8 B: transposed on the device (plumbing), so that every result column arrives
        # contiguous in a pinned buffer kept per device
        cols = res.view(torch.float64).view(n, 11).t().contiguous()
        key = device.index
        host = _PINNED_RESULTS.get(key)
        if host is None or host.numel() < 11 * n:
            host = _PINNED_RESULTS[key] = torch.empty(11 * n, dtype=torch.float64, pin_memory=True)
        host[: 11 * n].view(11, n).copy_(cols, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        c = host.numpy()[: 11 * n].reshape(11, n)
        tail = np.ascontiguousarray(c[9:11]).view(np.uint32).reshape(2, n, 2)   # (tp, fp), (fn, status)
        status = tail[1, :, 1].view(np.int32)
        if errors == "raise":
            zero = status == 2
            if zero.any():
                raise ZeroDivisionError(f"float division by zero (pair {int(np.flatnonzero(zero)[0])}: constant data)")
        # the ratios were formed on the device with the reference's float64 operations (metrics.py:25-152)
        names = ("ffi", "mad_reduction", "std_reduction", "flagged_fraction", "iou", "precision", "recall", "f1", "dice")
        out = {k: c[i].copy() for i, k in enumerate(names)}
        out["tp"] = tail[0, :, 0].astype(np.int64)
        out["fp"] = tail[0, :, 1].astype(np.int64)
        out["fn"] = tail[1, :, 0].astype(np.int64)
    return out
