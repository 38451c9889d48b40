"""Oracle (test infrastructure): NumPy restatement of `Preprocessor.create_dataset`.

Follows the reference step by step -- same order of operations, same NumPy/SciPy
primitives where the primitive decides the numerics (`np.nanmedian`, `np.median`,
`np.sqrt`, `np.log10`, `np.abs`, `np.angle`, `np.random.permutation`) -- so that
the result is bit-identical to the reference on the same host.  It is NOT used
by the product path.  Citations are `/root/reference/rfi_toolbox/...:line`.

Pinned by tests/test_oracle_vs_reference.py (live import) and tests/golden/.
"""

from __future__ import annotations

import numpy as np

try:  # same primitive the reference calls (preprocessor.py:14,129,699,736)
    from scipy import stats as _sp_stats
except Exception:  # pragma: no cover - scipy is in the image
    _sp_stats = None

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)  # preprocessor.py:779
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)  # preprocessor.py:780


# --------------------------------------------------------------------------- helpers
def mad_omit(values):
    """scipy.stats.median_abs_deviation(x, axis=None, nan_policy="omit"), scale 1.0
    (preprocessor.py:129, 699, 736): median(|v - median(v)|) over the non-NaN values."""
    v = np.asarray(values).ravel()
    v = v[~np.isnan(v)]
    if v.size == 0:
        return v.dtype.type(np.nan)
    centre = np.median(v)
    return np.median(np.abs(v - centre))


def _mad(values):
    if _sp_stats is not None:
        return _sp_stats.median_abs_deviation(values, axis=None, nan_policy="omit")
    return mad_omit(values)


def cabs_numpy_algorithm(z):
    """What `np.abs(complex64/128)` computes on x86 SIMD builds of NumPy 2.x:
    hi * sqrt(fma(r, r, 1)) with r = lo / hi, hi = max(|re|,|im|), lo = min(...).
    Used only by tests to prove the CUDA magnitude formula equals `np.abs`
    (the oracle itself calls `np.abs`, as the reference does: preprocessor.py:127,574)."""
    z = np.asarray(z)
    re = np.abs(z.real).astype(np.longdouble)
    im = np.abs(z.imag).astype(np.longdouble)
    ft = z.real.dtype.type
    hi = np.maximum(re, im)
    lo = np.minimum(re, im)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = (lo.astype(ft) / hi.astype(ft)).astype(ft)  # one rounded division in T
        # fma(r, r, 1): exact product + 1 rounded once.  longdouble (64-bit mantissa)
        # holds r*r exactly for float32 only; tests restrict this helper to complex64.
        t = (r.astype(np.longdouble) * r.astype(np.longdouble) + 1).astype(ft)
        out = (np.sqrt(t) * hi.astype(ft)).astype(ft)
    out = np.where(hi == 0, ft(0), out)
    return out


def effective_rotations(enable_augmentation, augmentation_rotations):
    """preprocessor.py:240, 436-444: >=4 -> 4 views, 2..3 -> 2 views, else 1."""
    if not enable_augmentation or augmentation_rotations <= 1:
        return 1
    return 4 if augmentation_rotations >= 4 else 2


def rotate_views(cube, n_rot, augment):
    """preprocessor.py:413-446 (augment) / :252 (plain flatten)."""
    out = []
    for bl in cube:
        for wf in bl:
            out.append(wf)
            if not augment:
                continue
            if n_rot >= 2:
                out.append(np.ascontiguousarray(wf[::-1, :]))
            if n_rot >= 4:
                out.append(wf.T)
                out.append(np.ascontiguousarray(wf.T[::-1, :]))
    return out


def tile(waterfall, p):
    """patchify + padding (preprocessor.py:22-42, 512-558): zero-pad bottom/right to a
    multiple of p, then non-overlapping p x p tiles, row-major over (row-block, col-block).
    Returns an array (nh*nw, p, p)."""
    rows, cols = waterfall.shape
    pr = (p - rows) if rows < p else (-rows) % p
    pc = (p - cols) if cols < p else (-cols) % p
    if pr or pc:
        waterfall = np.pad(waterfall, ((0, pr), (0, pc)), mode="constant", constant_values=0)
    rows, cols = waterfall.shape
    nh, nw = rows // p, cols // p
    t = np.ascontiguousarray(waterfall).reshape(nh, p, nw, p).transpose(0, 2, 1, 3)
    return np.ascontiguousarray(t).reshape(nh * nw, p, p)


def tile_all(views, p):
    shapes = [tuple(v.shape) for v in views]
    return np.concatenate([tile(v, p) for v in views], axis=0), shapes


# --------------------------------------------------------------------------- per-patch stages
def normalize_patches(patches):
    """preprocessor.py:646-670."""
    out = []
    for patch in patches:
        if np.iscomplexobj(patch):
            patch = np.abs(patch)
        with np.errstate(all="ignore"):
            m = np.nanmedian(patch)
        out.append(patch / m if m > 0 else patch)
    return np.array(out)


def stretch_patches(patches, stretch):
    """preprocessor.py:672-706: sqrt|x| or log10|x|; +-inf := MAD(finite), NaN kept."""
    if stretch == "SQRT":
        fn = np.sqrt
    elif stretch == "LOG10":
        fn = np.log10
    else:
        raise ValueError(f"Invalid stretch '{stretch}'. Use 'SQRT' or 'LOG10'")
    out = []
    for patch in patches:
        with np.errstate(all="ignore"):
            s = fn(np.abs(patch))
        fin = s[np.isfinite(s)]
        if fin.size > 0:
            s[np.isinf(s)] = _mad(fin)
        else:
            s[np.isinf(s)] = 0
        out.append(s)
    return np.array(out)


def mad_flags(patches, sigma, abs_complex=True):
    """preprocessor.py:708-745 / :114-136.  For complex patches the reference's Pool
    route takes |z| first (:126-127); SURVEY section 8-a0 fixes that as the semantics."""
    out = []
    for patch in patches:
        if abs_complex and np.iscomplexobj(patch):
            patch = np.abs(patch)
        with np.errstate(all="ignore"):
            d = _mad(patch)
            c = np.nanmedian(patch)
            hi = c + (d * sigma)
            lo = c - (d * sigma)
            out.append((patch > hi) | (patch < lo))
    return np.array(out, dtype=bool)


def _minmax(ch):
    with np.errstate(all="ignore"):
        lo, hi = np.nanmin(ch), np.nanmax(ch)
    if hi > lo:
        return (ch - lo) / (hi - lo)
    return np.zeros_like(ch)


def _gradient(log_amp):
    td = np.zeros_like(log_amp)
    fd = np.zeros_like(log_amp)
    td[1:, :] = np.diff(log_amp, axis=0)
    fd[:, 1:] = np.diff(log_amp, axis=1)
    return np.sqrt(td**2 + fd**2)


def extract_channels_real(patch):
    """preprocessor.py:608-644 -> (H, W, 3) [minmax(grad), minmax(log_amp), 0]."""
    with np.errstate(all="ignore"):
        log_amp = np.log10(np.abs(patch) + 1e-10)
        g = _gradient(log_amp)
        return np.stack([_minmax(g), _minmax(log_amp), np.zeros_like(log_amp)], axis=-1)


def extract_channels_complex(patch):
    """preprocessor.py:562-606 -> (H, W, 3) [minmax(grad), clip((L+3)/7), (phase+pi)/2pi]."""
    with np.errstate(all="ignore"):
        log_amp = np.log10(np.abs(patch) + 1e-10)
        phase = np.angle(patch)
        g = _gradient(log_amp)
        lo, hi = -3.0, 4.0
        ch1 = np.clip((log_amp - lo) / (hi - lo), 0, 1)
        ch2 = (phase + np.pi) / (2 * np.pi)
        return np.stack([_minmax(g), ch1, ch2], axis=-1)


def images_exact64(patches):
    """The channel extraction + ImageNet step (preprocessor.py:562-644, 765-783) evaluated in
    FLOAT64 on the given processed patches (their float32 / complex64 values taken as exact).

    Not the reference's arithmetic -- a yardstick for it: NumPy's float32 `log10`, `arctan2`
    and the four rounded steps of min-max + ImageNet normalisation sit within a few ulp of this
    chain, but the per-patch min-max divides that noise by (hi - lo), so the distance between
    ANY two correct float32 implementations is `noise / (hi - lo) / std`, not 1e-6 relative.
    Tests measure |reference - exact64| (NumPy's own noise) and require the CUDA path to stay
    within the same distance of it.  Returns float64 (N, H, W, 3)."""
    mean = IMAGENET_MEAN.astype(np.float64)
    std = IMAGENET_STD.astype(np.float64)
    # the reference subtracts / divides by the float32 constants
    out = []
    for patch in patches:
        with np.errstate(all="ignore"):
            if np.iscomplexobj(patch):
                amp = np.abs(patch).astype(np.float64)      # np.abs(complex64) is reproduced bit-exactly
                phase = np.arctan2(patch.imag.astype(np.float64), patch.real.astype(np.float64))
            else:
                amp, phase = np.abs(patch.astype(np.float64)), None
            log_amp = np.log10(amp + 1e-10)
            g = _gradient(log_amp)
            if phase is not None:
                ch1 = np.clip((log_amp + 3.0) / 7.0, 0, 1)
                ch2 = (phase + np.pi) / (2 * np.pi)
            else:
                ch1, ch2 = _minmax(log_amp), np.zeros_like(log_amp)
            img = np.stack([_minmax(g), ch1, ch2], axis=-1)
        out.append((img - mean) / std)
    return np.array(out, dtype=np.float64)


# --------------------------------------------------------------------------- container
class OracleDataset:
    """Duck-type of datasets/batched_dataset.py:10-45 holding NumPy arrays."""

    def __init__(self, images, labels, metadata):
        assert len(images) == len(labels)
        assert images.dtype == np.float32 and labels.dtype == np.uint8
        self.images, self.labels, self.metadata = images, labels, metadata

    def __len__(self):
        return len(self.images)

    def __getitem__(self, i):
        return {"image": self.images[i], "label": self.labels[i]}


# --------------------------------------------------------------------------- the pipeline
def create_dataset(
    data,
    flags=None,
    patch_size=128,
    stretch=None,
    flag_sigma=5,
    use_custom_flags=True,
    num_patches=None,
    normalize_before_stretch=True,
    normalize_after_stretch=False,
    num_workers=4,
    enable_augmentation=True,
    augmentation_rotations=4,
    inference_mode=False,
    return_intermediates=False,
):
    """Preprocessor(data, flags).create_dataset(...) -- preprocessor.py:175-411.

    Draws exactly one `np.random.permutation` from the global legacy RNG unless
    `inference_mode` (preprocessor.py:760).  `num_workers` is accepted and ignored
    (the Pool and sequential routes agree for real input; for complex input the
    Pool semantics are used, see `mad_flags`)."""
    data = np.asarray(data)
    if data.ndim == 3:
        data = data[np.newaxis, ...]  # :187-189 (flags are NOT reshaped -- quirk Q7)
    elif data.ndim != 4:
        raise ValueError(f"Data must be 3D or 4D, got shape {data.shape}")

    augment = bool(enable_augmentation and augmentation_rotations > 1)
    views = rotate_views(data, augmentation_rotations, augment)
    fviews = None
    if use_custom_flags and flags is not None:
        flags = np.asarray(flags)
        if flags.ndim != 4:
            raise ValueError("flags must be 4-D (baselines, pols, channels, times)")
        fviews = rotate_views(flags, augmentation_rotations, augment)

    p = patch_size
    original_shapes = None
    if views[0].shape[0] <= p and views[0].shape[1] <= p:  # :261-269
        patches = np.array(views)
        fpatches = np.array(fviews) if fviews is not None else None
    else:
        patches, original_shapes = tile_all(views, p)
        fpatches = tile_all(fviews, p)[0] if fviews is not None else None

    if not np.iscomplexobj(patches):  # :285-313
        if normalize_before_stretch:
            patches = normalize_patches(patches)
        if stretch:
            patches = stretch_patches(patches, stretch)
        if normalize_after_stretch:
            patches = normalize_patches(patches)

    if inference_mode:  # :317-334
        pflags = np.zeros(patches.shape[:3], dtype=np.uint8)
    elif fpatches is not None:
        pflags = fpatches
    else:
        pflags = mad_flags(patches, flag_sigma)

    inter = {"n_total": len(patches)}
    if return_intermediates:
        inter["processed"] = patches.copy()
        inter["flags_canonical"] = np.asarray(pflags).copy()

    order = np.arange(len(patches))
    if not inference_mode:
        keep = np.array([f.any() for f in pflags])  # :746-756
        if keep.any():
            patches, pflags, order = patches[keep], pflags[keep], order[keep]
        perm = np.random.permutation(len(patches))  # :758-763
        patches, pflags, order = patches[perm], pflags[perm], order[perm]
    if num_patches and num_patches < len(patches):  # :356-359
        patches, pflags, order = patches[:num_patches], pflags[:num_patches], order[:num_patches]
    inter["order"] = order  # canonical (pre-compaction) index of every output patch

    imgs = []
    for patch in patches:  # :366-380
        ch = extract_channels_complex(patch) if np.iscomplexobj(patch) else extract_channels_real(patch)
        imgs.append(ch.astype(np.float32))
    images = np.array(imgs, dtype=np.float32)
    if images.size:
        images = (images - IMAGENET_MEAN) / IMAGENET_STD  # :765-783
    else:
        images = images.reshape((0,) + tuple(patches.shape[1:3]) + (3,))
    labels = np.array(pflags, dtype=np.uint8)  # :386

    metadata = {  # :394-402
        "patch_size": patch_size,
        "stretch": stretch,
        "flag_sigma": flag_sigma,
        "normalize_before_stretch": normalize_before_stretch,
        "normalize_after_stretch": normalize_after_stretch,
        "augmentation_rotations": augmentation_rotations,
        "original_shapes": original_shapes,
    }
    ds = OracleDataset(images.astype(np.float32, copy=False), labels, metadata)
    if return_intermediates:
        return ds, inter
    return ds


def canonical_index_map(n_waterfalls, n_rot, nh, nw):
    """canonical (pre-compaction) patch index of rotation r of
    original tile (i, j) of waterfall w, for dims divisible by the patch size.
    Returns int64 array [n_waterfalls, n_rot, nh, nw]."""
    per = nh * nw
    i = np.arange(nh)[:, None]
    j = np.arange(nw)[None, :]
    out = np.empty((n_waterfalls, n_rot, nh, nw), dtype=np.int64)
    for w in range(n_waterfalls):
        base = w * n_rot * per
        out[w, 0] = base + i * nw + j
        if n_rot >= 2:
            out[w, 1] = base + per + (nh - 1 - i) * nw + j
        if n_rot >= 4:
            out[w, 2] = base + 2 * per + j * nh + i
            out[w, 3] = base + 3 * per + (nw - 1 - j) * nh + i
    return out
