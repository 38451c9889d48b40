"""Oracle (test infrastructure): NumPy restatement of GPUPreprocessor.create_raw_patches
(rfi_toolbox/preprocessing/preprocessor.py:784-972).

The reference's method cannot run whenever it patchifies: `_create_patches` returns ONE list
(:972) but the caller unpacks two values (:893, :896), so only the "whole waterfall" branch
(:885-890) executes upstream.  This port restates the evident intent -- the same tiling as
`patchify(waterfall, (P, P), step=P)` (:22-42, :966-970) for data and flags -- and is pinned against
the live reference on the branch that does run (tests/test_oracle_vs_reference.py).

`num_workers` selects between the two tilings the reference holds for that branch (:954-970): a
positive value (the default, 4) is the worker-pool route through `_patchify_single_waterfall`
(:46-112), which zero-pads bottom / right to multiples of the patch size (a dimension below the
patch size is padded up to it); 0 is the in-process `patchify(...)` loop, which drops remainders."""
from __future__ import annotations

import numpy as np


def create_raw_patches(data, flags=None, patch_size=256, remove_blank=True, num_patches=None, num_workers=4):
    """-> (list of complex (H, W) arrays, list of bool (H, W) arrays); draws from the global
    legacy RNG exactly as the reference does (`np.random.choice` if truncating, then one
    `np.random.permutation`)."""
    data = np.asarray(data)
    if data.ndim == 3:                      # :825-830
        data = data[np.newaxis, ...]
    elif data.ndim != 4:
        raise ValueError(f"Data must be 3D or 4D, got shape {data.shape}")
    if not np.iscomplexobj(data):           # :832-836
        raise ValueError("GPUPreprocessor requires complex data. Use standard Preprocessor for real-valued data.")
    waterfalls = [pol for baseline in data for pol in baseline]                     # :875
    if flags is not None:
        masks = [pol for baseline in np.asarray(flags) for pol in baseline]        # :879
    else:
        masks = [np.abs(w) > 0 for w in waterfalls]                                 # :881-883
    rows, cols = waterfalls[0].shape
    p = int(patch_size)
    if rows <= p and cols <= p:             # :885-890
        patches, pmasks = list(waterfalls), list(masks)
    else:                                   # :966-970 for both lists
        if num_workers and num_workers > 0:   # :80-101, zero pad (False for the masks)
            pr = (-rows) % p if rows >= p else p - rows
            pc = (-cols) % p if cols >= p else p - cols
            if pr or pc:
                waterfalls = [np.pad(w, ((0, pr), (0, pc)), mode="constant", constant_values=0) for w in waterfalls]
                masks = [np.pad(np.asarray(m), ((0, pr), (0, pc)), mode="constant", constant_values=0) for m in masks]
                rows, cols = rows + pr, cols + pc
        patches, pmasks = [], []
        for w, m in zip(waterfalls, masks):
            for i in range(rows // p):
                for j in range(cols // p):
                    patches.append(w[i * p:(i + 1) * p, j * p:(j + 1) * p])
                    pmasks.append(m[i * p:(i + 1) * p, j * p:(j + 1) * p])
    if remove_blank:                        # :904-913
        keep = [bool(np.asarray(m).any()) for m in pmasks]
        patches = [x for x, k in zip(patches, keep) if k]
        pmasks = [x for x, k in zip(pmasks, keep) if k]
    if num_patches and num_patches < len(patches):                                  # :918-922
        idx = np.random.choice(len(patches), num_patches, replace=False)
        patches = [patches[i] for i in idx]
        pmasks = [pmasks[i] for i in idx]
    idx = np.random.permutation(len(patches))                                       # :925-928
    return [patches[i] for i in idx], [pmasks[i] for i in idx]
