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
