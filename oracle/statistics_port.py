"""Oracle (test infrastructure): NumPy restatement of evaluation/statistics.py:10-97."""

from __future__ import annotations

import numpy as np


def compute_mad(data):
    """statistics.py:10-13."""
    m = np.median(data)
    return np.median(np.abs(data - m))


def compute_statistics(data, flags=None):
    """statistics.py:16-56: stats of |data| over all samples or the unflagged ones,
    computed in the data's own precision and returned as Python floats."""
    data = np.asarray(data)
    if np.iscomplexobj(data):
        data = np.abs(data)
    if flags is not None:
        clean = data[~flags]  # flags must be bool (uint8 would index) -- quirk kept
        frac = np.sum(flags) / flags.size
    else:
        clean = data.ravel()
        frac = 0.0
    if len(clean) == 0:
        return {"mean": np.nan, "median": np.nan, "std": np.nan, "mad": np.nan,
                "count": 0, "flagged_fraction": 1.0}
    return {
        "mean": float(np.mean(clean)),
        "median": float(np.median(clean)),
        "std": float(np.std(clean)),
        "mad": float(compute_mad(clean)),
        "count": len(clean),
        "flagged_fraction": float(frac),
    }


def compute_ffi(data, flags):
    """statistics.py:59-97."""
    before = compute_statistics(data, None)
    after = compute_statistics(data, flags)
    if np.isnan(after["mad"]) or np.isnan(after["std"]):
        return {"ffi": 0.0, "mad_reduction": 0.0, "std_reduction": 0.0, "flagged_fraction": 1.0}
    mad_red = 1.0 - (after["mad"] / before["mad"])
    std_red = 1.0 - (after["std"] / before["std"])
    pen = after["flagged_fraction"]
    ffi = (0.5 * mad_red + 0.5 * std_red) * (1.0 - 0.5 * pen)
    return {"ffi": float(ffi), "mad_reduction": float(mad_red),
            "std_reduction": float(std_red), "flagged_fraction": float(pen)}


def compute_calcquality(data, flags, reference_data=None):
    """statistics.py:100-193."""
    data = np.asarray(data)
    if np.iscomplexobj(data):
        data = np.abs(data)
    if reference_data is not None:
        reference_data = np.asarray(reference_data)
        if np.iscomplexobj(reference_data):
            reference_data = np.abs(reference_data)
        ref_stats = compute_statistics(reference_data, None)
        ref_data = reference_data.ravel()
    else:
        ref_stats = compute_statistics(data, None)
        ref_data = data.ravel()
    flag_stats = compute_statistics(data, flags)
    rmean, rstd = ref_stats["mean"], ref_stats["std"]
    fmean, fstd = flag_stats["mean"], flag_stats["std"]
    pflag = flag_stats["flagged_fraction"] * 100
    if np.isnan(fmean) or np.isnan(fstd) or rstd < 1e-10:
        return {"calcquality": np.inf, "sensitivity": np.inf, "mean_shift": np.inf, "std_shift": np.inf,
                "overflagging_penalty": np.inf, "flagged_pct": float(pflag), "components": {}}
    rmax = np.max(ref_data)
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
