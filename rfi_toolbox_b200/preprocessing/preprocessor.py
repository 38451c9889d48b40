"""Preprocessor -- drop-in for rfi_toolbox/preprocessing/preprocessor.py:139-783.

`Preprocessor(data, flags).create_dataset(...)` keeps the reference's signature, defaults,
metadata dict, error behaviour and RNG consumption (exactly one `np.random.permutation`
from the global legacy RNG unless `inference_mode`), and returns a `TorchDataset` whose
`.images (N, P, P, 3) float32` / `.labels (N, P, P) uint8` hold the same values in the same
order.  What changes is where the work happens:

  phase 1  `rfi_tile_stats`    per-tile median / MAD / thresholds / flag counts   (GPU)
  host     keep mask -> stable compaction -> np.random.permutation -> dest_slot[] (tiny)
  phase 2  `rfi_write_patches` rotate + tile + normalise + stretch + flag + 3-channel
                               + ImageNet normalise, each patch written once to its final
                               shuffled slot                                        (GPU)

Extensions, all keyword-only and off by default: `magnitude=True` treats complex input as
the reference treats `np.abs(data)` (the real branch, with |z| fused into the load);
`device=`; `pin=` for the host->device copy of NumPy input; `compute_dtype="float32"` takes
complex128 / float64 input (what the reference's loaders emit, io/ms_loader.py:202-238) in as
complex64 / float32 instead of following it in float64 as the reference does (preprocessor.py:376).
"""
from __future__ import annotations

import ctypes as C
import logging
import threading

import numpy as np
import torch

from .. import _native
from ..datasets.batched_dataset import TorchDataset
from ..utils.device import as_device_tensor, as_host_tensor, copy_stream, current_stream_ptr, require_cuda

logger = logging.getLogger(__name__)

_DTYPE_CODE = {
    torch.float32: _native.RFI_F32, torch.float64: _native.RFI_F64,
    torch.complex64: _native.RFI_C64, torch.complex128: _native.RFI_C128,
}
_INT_DTYPES = (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64, torch.bool)
_STRETCH_CODE = {None: _native.RFI_STRETCH_NONE, "SQRT": _native.RFI_STRETCH_SQRT,
                 "LOG10": _native.RFI_STRETCH_LOG10}


def patchify(array, patch_shape, step):
    """preprocessor.py:22-42: (H, W) -> (n_h, n_w, patch_h, patch_w) windows.

    Pure index arithmetic (a strided view made contiguous) -- kept for API parity and for
    the reference's own known-answer tests (tests/test_preprocessing.py:14-66); the CUDA
    path never materialises this array."""
    ph, pw = patch_shape
    a = np.asarray(array)
    nh = (a.shape[0] - ph) // step + 1
    nw = (a.shape[1] - pw) // step + 1
    s0, s1 = a.strides
    view = np.lib.stride_tricks.as_strided(a, (nh, nw, ph, pw), (s0 * step, s1 * step, s0, s1),
                                           writeable=False)
    return np.ascontiguousarray(view)


def canonical_index_map(n_waterfalls, n_rot, nh, nw):
    """Canonical (pre-compaction) patch index of rotation r of original tile (i, j) of
    waterfall w -- the order in which preprocessor.py:413-446 + :478-560 emit patches.
    int64 [n_waterfalls, n_rot, nh, nw]."""
    per = nh * nw
    i = np.arange(nh, dtype=np.int64)[:, None]
    j = np.arange(nw, dtype=np.int64)[None, :]
    local = np.empty((n_rot, nh, nw), dtype=np.int64)
    local[0] = i * nw + j
    if n_rot >= 2:
        local[1] = per + (nh - 1 - i) * nw + j
    if n_rot >= 4:
        local[2] = 2 * per + j * nh + i
        local[3] = 3 * per + (nw - 1 - j) * nh + i
    base = np.arange(n_waterfalls, dtype=np.int64)[:, None, None, None] * (n_rot * per)
    return base + local[None]


def _keep_in_canonical_order(keep_tile, n_rot):
    """Per-tile keep mask (W, nh, nw) -> per-patch mask in the reference's patch order: the
    rotated waterfalls of w are tiled row-major, so rotation 1 lists the tiles with the row
    blocks reversed, rotation 2 lists the transposed grid, rotation 3 the transposed grid with
    its row blocks reversed (same map as `canonical_index_map`, without building indices)."""
    w = keep_tile.shape[0]
    parts = [keep_tile.reshape(w, -1)]
    if n_rot >= 2:
        parts.append(keep_tile[:, ::-1, :].reshape(w, -1))
    if n_rot >= 4:
        t = keep_tile.transpose(0, 2, 1)
        parts.append(t.reshape(w, -1))
        parts.append(t[:, ::-1, :].reshape(w, -1))
    return np.concatenate(parts, axis=1).reshape(-1)


class _HostBuffers:
    """Pinned staging buffers of ONE call in flight (flag counts down, destination slots up)."""

    def __init__(self, n_groups, n_patches):
        self.nflag = torch.empty(n_groups, dtype=torch.int32, pin_memory=True)
        self.dest = torch.empty(n_patches, dtype=torch.int64, pin_memory=True)
        self.order = torch.empty(n_patches, dtype=torch.int64)
        self.nflag_np, self.dest_np, self.order_np = self.nflag.numpy(), self.dest.numpy(), self.order.numpy()
        self.event = torch.cuda.Event()    # the flag counts have landed
        self.copied = torch.cuda.Event()   # the destination slots have left the pinned buffer
        self.busy = False


_HOST_POOL = {}
_HOST_POOL_LOCK = threading.Lock()

# ---- single-launch path: bookkeeping of the speculative shuffle draws (see `Preprocessor.speculate`)
_SPEC_LOCK = threading.Lock()
_SPEC_INFLIGHT = []      # contexts of speculative calls whose result() has not run yet, in submission order
_SPEC_COOLDOWN = {}      # call signature -> submissions left before speculation is tried again
_SPEC_BACKOFF = {}       # call signature -> length of its last cool-down (doubles per failure)
_ZERO_COUNT = np.zeros(1, dtype=np.int32)


def _acquire_host_buffers(device, n_groups, n_patches):
    """Pinned memory is expensive to allocate (more than the whole host phase), so the staging
    buffers are pooled per device and grown on demand.  Every call in flight -- several when
    `create_dataset_async` is used with lookahead, or from several threads -- owns its own set
    from `acquire` until its destination slots have been uploaded (`_release_host_buffers`)."""
    with _HOST_POOL_LOCK:
        pool = _HOST_POOL.setdefault(device.index, [])
        for hb in pool:
            if not hb.busy and hb.nflag.numel() >= n_groups and hb.dest.numel() >= n_patches:
                hb.busy = True
                break
        else:
            big = max(pool, key=lambda h: h.dest.numel(), default=None)
            g = max(n_groups, 2 * big.nflag.numel() if big and big.nflag.numel() < n_groups else 0)
            n = max(n_patches, 2 * big.dest.numel() if big and big.dest.numel() < n_patches else 0)
            hb = _HostBuffers(g, n)
            hb.busy = True
            pool[:] = [h for h in pool if h.busy or (h.nflag.numel() >= g and h.dest.numel() >= n)][-7:] + [hb]
    hb.copied.synchronize()  # a previous owner's upload has left the pinned buffer
    return hb


def _release_host_buffers(hb):
    with _HOST_POOL_LOCK:
        hb.busy = False


class PendingDataset:
    """Handle of a `create_dataset_async` call: phase 1 (statistics and flag counts) is enqueued, the
    host phase and phase 2 run in `result()`.

    `result()` draws the call's ONE `np.random.permutation` from the global legacy generator
    (preprocessor.py:760), so call it in the order the reference would have called
    `create_dataset` -- the k-th `result()` then consumes exactly the k-th call's share of the
    stream.  Between `create_dataset_async` and `result()` the GPU works on this call's statistics
    while the host (and the GPU) finish the previous call: the host phase of call k -- waiting for
    the flag counts, the shuffle, the upload of the destination slots -- hides behind phase 1 of
    call k + 1."""

    def __init__(self, pre, ctx=None, dataset=None):
        self._pre, self._ctx, self._dataset = pre, ctx, dataset

    def done(self):
        return self._dataset is not None

    def result(self):
        if self._dataset is None:
            self._dataset = self._pre._complete(self._ctx)
            self._ctx = None
        return self._dataset

    def __del__(self):
        # a handle dropped without result(): its staging buffers go back to the pool and it no longer
        # counts as a speculative call in flight (its shuffle draw stays consumed, like a discarded dataset's)
        ctx = getattr(self, "_ctx", None)
        if ctx is not None:
            try:
                with _SPEC_LOCK:
                    _SPEC_INFLIGHT[:] = [c for c in _SPEC_INFLIGHT if c is not ctx]
                _release_host_buffers(ctx["hb"])
            except Exception:
                pass


class Preprocessor:
    """Preprocess waterfall data into training patches (preprocessor.py:139-196)."""

    #: host (pinned) input is uploaded in this many baseline chunks, overlapped with phase 1
    upload_chunks = 8

    #: when True, CUDA events bracket the two kernels of every call (read by bench.py):
    #: `self.events = {"stats": (start, stop), "write": (start, stop)}` on the current stream
    #: (`{"fused": (start, stop)}` when the call took the single-launch path).
    profile = False

    #: Single-launch path (`rfi_fused_patches`: statistics and patches of a tile in ONE kernel, no host
    #: round trip between the phases), OFF by default: measured on B200 it moves fewer bytes (60 instead of
    #: 68 B / px) but takes 3.85 ms where the two launches take 2.87 ms on the bench cube -- a CTA that holds
    #: both halves runs them one after the other, and statistics and writer CTAs side by side on one SM
    #: evict each other's instructions.  It remains for callers that cannot synchronise with
    #: the host between the phases (stream capture, `inference_mode` services).
    #: The kernel needs every patch's destination slot before the statistics exist.  `inference_mode` has
    #: them by definition.  With MAD flags the call SPECULATES that no patch is blank: it draws
    #: `np.random.permutation(N0)` at submission, launches, and checks the flag counts in `result()`.  If a
    #: tile did turn out blank (while others were not), the global generator is put back to where it stood
    #: before the draw and the call completes through the two-phase path with the statistics the launch
    #: left -- results and RNG consumption are the reference's either way.  Calls submitted after a
    #: mis-speculated one and before its `result()` are re-drawn in order; a call signature that
    #: mis-speculated is not tried again for a while.
    speculate = False

    def __init__(self, data, flags=None, *, magnitude=False, device=None, pin=False, compute_dtype=None):
        ndim = data.ndim
        if ndim == 3:
            data = data[None, ...]  # :187-189 -- flags are NOT reshaped (reference quirk Q7)
        elif ndim != 4:
            raise ValueError(f"Data must be 3D or 4D, got shape {tuple(data.shape)}")
        self.data = data
        self.flags = flags
        self._patches, self._patch_src = None, None
        self.patch_flags = None
        self.dataset = None
        self.magnitude = bool(magnitude)
        self._device = device
        self._pin = pin
        if compute_dtype not in (None, "float32", "float64"):
            raise ValueError("compute_dtype must be None (follow the input, as the reference does), 'float32' or 'float64'")
        self._compute_dtype = compute_dtype
        self.last_tile_stats = None
        self.last_launch = None   # "single" / "two-phase" / ... : which launches the last fast-path call made

    # ------------------------------------------------------------------------------ helpers
    @staticmethod
    def _effective_rotations(enable_augmentation, augmentation_rotations):
        """preprocessor.py:240, 436-444."""
        if not enable_augmentation or augmentation_rotations <= 1:
            return 1
        return 4 if augmentation_rotations >= 4 else 2

    def _ingest_float32(self, lib, device, data):
        """complex128 -> complex64 / float64 -> float32 on the device, one baseline at a time, so that at
        most one baseline of the wide type is resident beside the narrow cube."""
        narrow = torch.complex64 if data.is_complex() else torch.float32
        with torch.cuda.device(device):
            out = torch.empty(data.shape, dtype=narrow, device=device)
            stream = current_stream_ptr(device)
            code = _DTYPE_CODE[data.dtype]
            for b in range(data.shape[0]):
                wide = data[b]
                if wide.device != device:
                    wide = wide.to(device, non_blocking=True)
                wide = wide.contiguous()
                _native.check(lib.rfi_downcast(wide.data_ptr(), out[b].data_ptr(), code, wide.numel(), stream), "rfi_downcast")
        return out

    @property
    def patches(self):
        """preprocessor.py:194, 272-311, 345-359: the processed patches of the last `create_dataset` call in
        the dataset's order -- `(N, P, P)` in the data's precision on the real branch (normalised, stretched,
        inf-filled), the raw complex samples on the complex branch.  The hot path goes from the cube straight
        to the image channels, so the array is built on first access (`rfi_processed_patches`) from the
        input, the tile statistics and the patch order the call left, and cached.  None before the first call."""
        if self._patches is None and self._patch_src is not None:
            lib = _native.load()
            plan, stats, order, device = self._patch_src
            if lib.rfi_plan_path(C.byref(plan)) == _native.RFI_PATH_GENERIC:
                raise NotImplementedError("Preprocessor.patches is rebuilt only for waterfalls whose dims are multiples of "
                                          "the patch size (P = 128 / 256 / 512 / 1024)")
            data = as_device_tensor(self.data, device)
            if self._compute_dtype == "float32" and data.dtype in (torch.float64, torch.complex128):
                data = self._ingest_float32(lib, device, data)
            cb = data.is_complex() and not self.magnitude
            real_t = torch.float32 if data.dtype in (torch.float32, torch.complex64) else torch.float64
            with torch.cuda.device(device):
                out = torch.empty((len(order), plan.patch, plan.patch), dtype=data.dtype if cb else real_t, device=device)
                od = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int64)).to(device)
                _native.check(lib.rfi_processed_patches(C.byref(plan), data.data_ptr(), stats.data_ptr(), od.data_ptr(),
                                                        len(order), out.data_ptr(), current_stream_ptr(device)),
                              "rfi_processed_patches")
            self._patches = out
        return self._patches

    @patches.setter
    def patches(self, value):
        self._patches = value

    def _resolve_device(self):
        for x in (self.data, self.flags):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                return require_cuda(x.device)
        return require_cuda(self._device)

    # ------------------------------------------------------------------------------ main
    def create_dataset(
        self,
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
    ):
        """preprocessor.py:198-411.  `num_workers` is accepted and ignored (there is no
        process pool on the GPU path)."""
        return self.create_dataset_async(
            patch_size=patch_size, stretch=stretch, flag_sigma=flag_sigma, use_custom_flags=use_custom_flags,
            num_patches=num_patches, normalize_before_stretch=normalize_before_stretch,
            normalize_after_stretch=normalize_after_stretch, num_workers=num_workers,
            enable_augmentation=enable_augmentation, augmentation_rotations=augmentation_rotations,
            inference_mode=inference_mode).result()

    def create_dataset_async(
        self,
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
        *,
        input_ready=False,
    ):
        """Same arguments as `create_dataset`; enqueues phase 1 and returns a `PendingDataset` whose
        `result()` is the dataset.  `create_dataset(...)` == `create_dataset_async(...).result()`.
        Streaming callers (one Preprocessor per sample / baseline chunk, the reference's own unit of
        work, synthetic_generator.py:55-107) keep calls in flight so that the host phase of each hides
        behind the statistics kernel of the next (`iter_dataset_chunks(lookahead=)`).
        `input_ready` is accepted for compatibility and ignored."""
        lib = _native.load()
        device = self._resolve_device()
        if stretch and stretch not in ("SQRT", "LOG10"):
            raise ValueError(f"Invalid stretch '{stretch}'. Use 'SQRT' or 'LOG10'")

        # host input: uploaded below in baseline chunks on a side stream, each chunk's statistics
        # kernel starting as soon as its chunk has landed (fast path); else one plain H2D copy
        host = as_host_tensor(self.data, pin=self._pin)
        data = host if host is not None else as_device_tensor(self.data, device, pin=self._pin)
        if data.dtype in _INT_DTYPES:
            # integer samples: NumPy promotes them to float64 at the first division / sqrt / log10 of the
            # reference chain, and every later step runs in float64 -- same results as float64 input
            data = data.to(torch.float64)
            if host is not None:
                host = data
        if data.dtype not in _DTYPE_CODE:
            raise TypeError(f"unsupported data dtype {data.dtype}: float32/64, complex64/128 or an integer type")
        if self._compute_dtype == "float32" and data.dtype in (torch.float64, torch.complex128):
            # loader-shaped complex128 / float64 input (io/ms_loader.py:202-238) taken in as complex64 /
            # float32: uploaded (if on the host) and rounded baseline by baseline (`rfi_downcast`)
            data, host = self._ingest_float32(lib, device, data), None
        B, npol, C_, T_ = data.shape
        is_complex = data.is_complex()
        complex_branch = is_complex and not self.magnitude

        custom = bool(use_custom_flags and self.flags is not None)
        flags = None
        if custom:
            if self.flags.ndim != 4:
                # the reference fails unpacking `.shape` of a 1-D row here (quirk Q7)
                raise ValueError("flags must be 4-D (baselines, pols, channels, times)")
            if tuple(self.flags.shape) != tuple(data.shape):
                raise ValueError(f"flags shape {tuple(self.flags.shape)} != data shape {tuple(data.shape)}")
        if custom and not inference_mode:
            flags = as_device_tensor(self.flags, device, pin=self._pin)
            if flags.dtype in (torch.bool, torch.uint8, torch.int8):
                flags = flags.view(torch.uint8)
            elif not flags.dtype.is_complex:
                # labels are the reference's `np.array(patch_flags, dtype=np.uint8)` (:386): other integer and
                # floating flag arrays (0 / 1 masks kept as int64 or float32) are cast the same way, once, on the
                # device.  (Blank patches are found on the cast bytes: a value that is a non-zero multiple of 256,
                # or a fraction inside (-1, 1), counts as unflagged there, while the reference's `.any()` sees it.)
                flags = flags.to(torch.uint8)
            else:
                raise TypeError(f"unsupported flags dtype {flags.dtype}")

        R = self._effective_rotations(enable_augmentation, augmentation_rotations)
        P = int(patch_size)
        skip_patchify = C_ <= P and T_ <= P  # :261
        if skip_patchify:
            if R == 4 and C_ != T_:
                raise ValueError("setting an array element with a sequence: rotated views of a "
                                 "non-square waterfall cannot be stacked (reference raises here too)")
        else:
            # :282 -- one (rows, cols) entry per rotated waterfall
            shapes = []
            for _ in range(B * npol):
                shapes.append((C_, T_))
                if R >= 2:
                    shapes.append((C_, T_))
                if R >= 4:
                    shapes.extend([(T_, C_), (T_, C_)])
            self.original_shapes = shapes

        if inference_mode:
            flag_mode = _native.RFI_FLAGS_INFERENCE
        elif custom:
            flag_mode = _native.RFI_FLAGS_CUSTOM
        else:
            flag_mode = _native.RFI_FLAGS_MAD

        plan = _native.RfiPlan(
            dtype=_DTYPE_CODE[data.dtype], magnitude=int(self.magnitude and is_complex),
            n_waterfalls=B * npol, channels=C_, times=T_, patch=P, rotations=R,
            stretch=_STRETCH_CODE[stretch] if not complex_branch else _native.RFI_STRETCH_NONE,
            norm_before=int(bool(normalize_before_stretch) and not complex_branch),
            norm_after=int(bool(normalize_after_stretch) and not complex_branch),
            flag_mode=flag_mode, sigma=float(flag_sigma),
        )
        # statistic groups: one per original tile (shared by its R rotations) when the dims are
        # multiples of P, one per output patch when the reference pads (the pad follows the flip)
        n_tiles = int(lib.rfi_plan_num_tiles(C.byref(plan)))
        n0 = int(lib.rfi_plan_num_patches(C.byref(plan)))
        if n_tiles < 0 or n0 < 0:
            _native.check(_native.RFI_E_INVALID, "rfi_plan_num_tiles")
        padded = (not skip_patchify) and (C_ % P != 0 or T_ % P != 0)
        nh, nw = (1, 1) if skip_patchify else (-(-C_ // P), -(-T_ // P))
        ws_bytes = int(lib.rfi_plan_workspace_bytes(C.byref(plan)))

        views = self._view_plans(lib, plan, R, C_, T_, P) if padded else None
        if views is not None:
            if host is not None:
                data = host.to(device, non_blocking=True)
            images, labels, order = self._create_padded(lib, device, plan, views, data, flags, R, B * npol, C_, T_, P,
                                                        inference_mode, num_patches)
            return PendingDataset(self, dataset=self._finish(
                images, labels, order, patch_size, stretch, flag_sigma, normalize_before_stretch,
                normalize_after_stretch, augmentation_rotations, patch_src=(plan, None, order, device)))

        # ---- single launch?  (inference_mode, or MAD flags with the all-kept shuffle drawn ahead)
        sig = (P, stretch, float(flag_sigma), bool(normalize_before_stretch), bool(normalize_after_stretch), R)
        fused = bool(self.speculate) and n_tiles > 0 and bool(lib.rfi_plan_fusable(C.byref(plan)))
        if fused and not inference_mode:
            with _SPEC_LOCK:
                left = _SPEC_COOLDOWN.get(sig, 0)
                if left > 0:
                    _SPEC_COOLDOWN[sig] = left - 1
                    fused = False
                elif any(c.get("stale") for c in _SPEC_INFLIGHT):
                    fused = False  # a re-draw is pending: later calls must draw after it, i.e. in result()

        with torch.cuda.device(device):
            stats = torch.empty((max(n_tiles, 1), _native.TILE_STAT_BYTES), dtype=torch.uint8, device=device)
            hb = _acquire_host_buffers(device, max(n_tiles, 1), max(n0, 1))
            fast = lib.rfi_plan_path(C.byref(plan)) == _native.RFI_PATH_FAST
        ctx = dict(pre=self, lib=lib, device=device, plan=plan, host=host, data=data, flags=flags, stats=stats, work=None, hb=hb,
                   ev=[torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.profile else None,
                   n_tiles=n_tiles, n0=n0, P=P, C_=C_, T_=T_, B=B, npol=npol, R=R, ws_bytes=ws_bytes, fast=fast,
                   skip_patchify=skip_patchify, inference_mode=inference_mode, num_patches=num_patches,
                   keep_work=not (fast and not (is_complex and self.magnitude)), fused=fused, sig=sig,
                   spec=None, stale=False,
                   meta=(patch_size, stretch, flag_sigma, normalize_before_stretch, normalize_after_stretch,
                         augmentation_rotations))
        self._launch_phase1(ctx)
        return PendingDataset(self, ctx=ctx)

    def _launch_phase1(self, ctx):
        """Enqueue phase 1 of `ctx` (the single launch when `ctx["fused"]`), then the download of its flag counts."""
        lib, device, plan, hb, ev = ctx["lib"], ctx["device"], ctx["plan"], ctx["hb"], ctx["ev"]
        host, data, flags, stats = ctx["host"], ctx["data"], ctx["flags"], ctx["stats"]
        n_tiles, n0, P, R, B, npol = ctx["n_tiles"], ctx["n0"], ctx["P"], ctx["R"], ctx["B"], ctx["npol"]
        C_, T_, ws_bytes, fast, fused = ctx["C_"], ctx["T_"], ctx["ws_bytes"], ctx["fast"], ctx["fused"]
        inference_mode, num_patches = ctx["inference_mode"], ctx["num_patches"]
        with torch.cuda.device(device):
            stream = current_stream_ptr(device)
            fptr = flags.data_ptr() if flags is not None else None
            if fused:
                # destination slots of the all-kept case (every group "unflagged" = nothing dropped,
                # preprocessor.py:752-756), the call's ONE draw from the global legacy generator
                rng_before = None if inference_mode else np.random.get_state()
                n_out = _native.plan_slots(plan, None if inference_mode else np.broadcast_to(_ZERO_COUNT, (n_tiles,)),
                                           not inference_mode, num_patches, hb.order_np, hb.dest_np)
                dest_dev = torch.empty(max(n0, 1), dtype=torch.int64, device=device)
                dest_dev.copy_(hb.dest[:max(n0, 1)], non_blocking=True)
                hb.copied.record()
                images = torch.empty((n_out, P, P, 3), dtype=torch.float32, device=device)
                labels = torch.empty((n_out, P, P), dtype=torch.uint8, device=device)
                ctx["spec"] = dict(rng_before=rng_before, n_out=n_out, images=images, labels=labels,
                                   order=hb.order_np[:n_out].copy(), sig=ctx["sig"])
                ws_bytes = 0
            if ev:
                ev[0].record()
            # ---- phase 1: statistics + flag counts per original tile (single launch: and the patches)
            work = torch.empty(ws_bytes, dtype=torch.uint8, device=device) if ws_bytes else None
            wptr = work.data_ptr() if work is not None else None

            def launch(pl, dptr, fp, sptr, wp, slot_off):
                if fused:
                    rc = lib.rfi_fused_patches(C.byref(pl), dptr, fp, sptr, dest_dev.data_ptr() + 8 * slot_off,
                                               images.data_ptr(), labels.data_ptr(), stream)
                    _native.check(rc, "rfi_fused_patches")
                else:
                    rc = lib.rfi_tile_stats(C.byref(pl), dptr, fp, sptr, wp, stream)
                    _native.check(rc, "rfi_tile_stats")

            pipelined = host is not None and host.is_pinned() and fast and B > 1 and n_tiles > 0
            if host is not None and not pipelined:
                data = host.to(device, non_blocking=True)
            if pipelined:
                # waterfalls are independent in phase 1: chunk c's kernel only waits for chunk c's copy
                data = torch.empty(host.shape, dtype=host.dtype, device=device)
                main, side = torch.cuda.current_stream(device), copy_stream(device)
                side.wait_stream(main)  # the fresh buffer may still be in use by work queued on `main`
                bounds = np.linspace(0, B, min(B, ctx["pre"].upload_chunks) + 1).astype(np.int64)
                per_bl_tiles = n_tiles // B
                for b0, b1 in zip(bounds[:-1], bounds[1:]):
                    b0, b1 = int(b0), int(b1)
                    with torch.cuda.stream(side):
                        data[b0:b1].copy_(host[b0:b1], non_blocking=True)
                        landed = torch.cuda.Event()
                        landed.record(side)
                    main.wait_event(landed)
                    sub = _native.RfiPlan.from_buffer_copy(plan)
                    sub.n_waterfalls = (b1 - b0) * npol
                    off = b0 * npol * C_ * T_
                    launch(sub, data.data_ptr() + off * data.element_size(), fptr + off if fptr is not None else None,
                           stats.data_ptr() + b0 * per_bl_tiles * _native.TILE_STAT_BYTES,
                           wptr + b0 * per_bl_tiles * (ws_bytes // n_tiles) if wptr else None,
                           b0 * per_bl_tiles * R)
                data.record_stream(side)
            else:
                launch(plan, data.data_ptr(), fptr, stats.data_ptr(), wptr, 0)
            if ev:
                ev[1].record()
            ctx["pre"].last_tile_stats = stats
            # real input: the fast path's scratch serves phase 1 only -- its block goes back before the
            # outputs are allocated (stream-ordered reuse).  Complex input through the real branch keeps it:
            # phase 2 reads the exact magnitudes phase 1 left there.
            ctx["data"], ctx["work"] = data, (work if ctx["keep_work"] else None)
            # ---- flag counts down to pinned memory; the host phase (`_complete`) waits for them
            if not inference_mode:
                hb.nflag[:n_tiles].copy_(stats[:n_tiles].view(torch.int32)[:, 16], non_blocking=True)
                hb.event.record()
        if ctx["spec"] is not None and not inference_mode:
            with _SPEC_LOCK:
                _SPEC_INFLIGHT.append(ctx)

    def _complete(self, ctx):
        """Host phase + phase 2 of a call whose phase 1 is enqueued (see `PendingDataset`); for a call
        on the single-launch path: the check of the speculation, and the two-phase completion if it failed."""
        lib, device, plan, hb, ev = ctx["lib"], ctx["device"], ctx["plan"], ctx["hb"], ctx["ev"]
        n_tiles, n0, P, C_, T_ = ctx["n_tiles"], ctx["n0"], ctx["P"], ctx["C_"], ctx["T_"]
        skip_patchify, inference_mode, num_patches = ctx["skip_patchify"], ctx["inference_mode"], ctx["num_patches"]
        data, flags, stats, work = ctx["data"], ctx["flags"], ctx["stats"], ctx["work"]
        spec = ctx["spec"]
        if spec is not None:
            good = True
            nflag = None
            if not inference_mode:
                hb.event.synchronize()
                nflag = hb.nflag_np[:n_tiles]
                any_flag = nflag > 0
                n_any = int(np.count_nonzero(any_flag))
                with _SPEC_LOCK:
                    _SPEC_INFLIGHT[:] = [c for c in _SPEC_INFLIGHT if c is not ctx]
                    if ctx["stale"]:
                        good = False   # an earlier call was re-drawn: this call's draw comes now, in order
                    elif n_any not in (0, n_tiles):
                        good = False   # a blank tile among flagged ones: the drawn permutation is too long
                        np.random.set_state(spec["rng_before"])
                        for c in _SPEC_INFLIGHT:
                            c["stale"] = True
                        # not tried again for a while, longer after every failure
                        _SPEC_COOLDOWN[spec["sig"]] = _SPEC_BACKOFF[spec["sig"]] = min(4096, 2 * _SPEC_BACKOFF.get(spec["sig"], 32))
                        logger.info("single-launch path: %d of %d tiles blank, completing through two phases",
                                    n_tiles - n_any, n_tiles)
            self.last_launch = "single" if good else "single, then phase 2 again (speculation failed)"
            if good:
                if ev:
                    self.events = {"fused": (ev[0], ev[1])}
                if nflag is not None and n_any == 0:
                    logger.warning("No flagged patches found - keeping all patches")
                _release_host_buffers(hb)
                return self._finish(spec["images"], spec["labels"], spec["order"], *ctx["meta"],
                                    patch_src=(plan, ctx["stats"], spec["order"], device))
            spec["images"] = spec["labels"] = None  # released before the right-sized outputs are allocated
        else:
            self.last_launch = "two-phase"
        with torch.cuda.device(device):
            stream = current_stream_ptr(device)
            fptr = flags.data_ptr() if flags is not None else None
            wptr = work.data_ptr() if work is not None else None
            # ---- host: blank-patch compaction + shuffle -> destination slot of every patch
            #      (one native call on pinned buffers; draws the ONE np.random.permutation of
            #      preprocessor.py:760 from the global legacy stream)
            nflag = None
            if not inference_mode:
                hb.event.synchronize()
                nflag = hb.nflag_np[:n_tiles]
            n_out = _native.plan_slots(plan, nflag, not inference_mode, num_patches, hb.order_np, hb.dest_np)
            dest_dev = torch.empty(max(n0, 1), dtype=torch.int64, device=device)
            dest_dev.copy_(hb.dest[:max(n0, 1)], non_blocking=True)
            hb.copied.record()

            # ---- phase 2: every kept patch written once, at its final position
            images = torch.empty((n_out, P if not skip_patchify else C_, P if not skip_patchify else T_, 3),
                                 dtype=torch.float32, device=device)
            labels = torch.empty(images.shape[:3], dtype=torch.uint8, device=device)
            if ev:
                ev[2].record()
            rc = lib.rfi_write_patches(C.byref(plan), data.data_ptr(), fptr, stats.data_ptr(),
                                       dest_dev.data_ptr(), images.data_ptr(), labels.data_ptr(), wptr, stream)
            _native.check(rc, "rfi_write_patches")
            if ev:
                ev[3].record()
                self.events = {"stats": (ev[0], ev[1]), "write": (ev[2], ev[3])}
            # host bookkeeping after the launch: the GPU idles between the phases, not here
            if not inference_mode and nflag is not None and not (nflag > 0).any():
                logger.warning("No flagged patches found - keeping all patches")
            order = hb.order_np[:n_out].copy()
            _release_host_buffers(hb)

        return self._finish(images, labels, order, *ctx["meta"], patch_src=(plan, stats, order, device))

    # ---------------------------------------------------------------- zero-padded geometries
    @staticmethod
    def _view_plans(lib, plan, R, C_, T_, P):
        """Dims that are not multiples of P: the reference pads every ROTATED view bottom / right
        (preprocessor.py:527-550), so each rotated patch is its own statistics group.  If a
        single-view plan over the padded view takes one of the on-chip paths, run the R views as
        R such plans over rotated, zero-padded copies (`rfi_rotate_pad`) instead of the generic
        path.  -> list of (plan_r, rows, cols) or None."""
        Cp, Tp = -(-C_ // P) * P, -(-T_ // P) * P
        out = []
        for r in range(R):
            rows, cols = (Cp, Tp) if r <= 1 else (Tp, Cp)
            pr = _native.RfiPlan.from_buffer_copy(plan)
            pr.channels, pr.times, pr.rotations = rows, cols, 1
            if lib.rfi_plan_path(C.byref(pr)) == _native.RFI_PATH_GENERIC:
                return None
            out.append((pr, rows, cols))
        return out

    def _create_padded(self, lib, device, plan, views, data, flags, R, n_wf, C_, T_, P, inference_mode, num_patches):
        per = (-(-C_ // P)) * (-(-T_ // P))
        n0 = n_wf * R * per
        with torch.cuda.device(device):
            stream = current_stream_ptr(device)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if self.profile else None
            if ev:
                ev[0].record()
            state = []
            for r, (pr, rows, cols) in enumerate(views):
                def rotate(src, esize, dtype):
                    dst = torch.empty((n_wf, rows, cols), dtype=dtype, device=device)
                    for w0 in range(0, n_wf, 65535):  # grid.z limit of one launch
                        nw_ = min(65535, n_wf - w0)
                        _native.check(lib.rfi_rotate_pad(src.data_ptr() + w0 * C_ * T_ * esize,
                                                         dst.data_ptr() + w0 * rows * cols * esize, esize, nw_, C_, T_,
                                                         rows, cols, r, stream), "rfi_rotate_pad")
                    return dst
                rot = rotate(data, data.element_size(), data.dtype)
                frot = rotate(flags, 1, torch.uint8) if flags is not None else None
                stats = torch.empty((n_wf * per, _native.TILE_STAT_BYTES), dtype=torch.uint8, device=device)
                ws = int(lib.rfi_plan_workspace_bytes(C.byref(pr)))
                work = torch.empty(ws, dtype=torch.uint8, device=device) if ws else None
                _native.check(lib.rfi_tile_stats(C.byref(pr), rot.data_ptr(), frot.data_ptr() if frot is not None else None,
                                                 stats.data_ptr(), work.data_ptr() if work is not None else None, stream),
                              "rfi_tile_stats")
                state.append((pr, rot, frot, stats, work))
            if ev:
                ev[1].record()
            # canonical order of a padded geometry: [waterfall][view][tile of the rotated grid]
            hb = _acquire_host_buffers(device, max(n0, 1), max(n0, 1))
            nflag = None
            if not inference_mode:
                counts = torch.stack([st[3].view(torch.int32)[:, 16].reshape(n_wf, per) for st in state], dim=1)
                hb.nflag[:n0].copy_(counts.reshape(-1), non_blocking=True)
                hb.event.record()
                hb.event.synchronize()
                nflag = hb.nflag_np[:n0]
            n_out = _native.plan_slots(plan, nflag, not inference_mode, num_patches, hb.order_np, hb.dest_np)
            dest = hb.dest[:n0].to(device, non_blocking=True).view(n_wf, R, per)
            hb.copied.record()
            images = torch.empty((n_out, P, P, 3), dtype=torch.float32, device=device)
            labels = torch.empty((n_out, P, P), dtype=torch.uint8, device=device)
            if ev:
                ev[2].record()
            for r, (pr, rot, frot, stats, work) in enumerate(state):
                dest_r = dest[:, r, :].contiguous()
                _native.check(lib.rfi_write_patches(C.byref(pr), rot.data_ptr(), frot.data_ptr() if frot is not None else None,
                                                    stats.data_ptr(), dest_r.data_ptr(), images.data_ptr(), labels.data_ptr(),
                                                    work.data_ptr() if work is not None else None, stream),
                              "rfi_write_patches")
            if ev:
                ev[3].record()
                self.events = {"stats": (ev[0], ev[1]), "write": (ev[2], ev[3])}
            if not inference_mode and nflag is not None and not (nflag > 0).any():
                logger.warning("No flagged patches found - keeping all patches")
            order = hb.order_np[:n_out].copy()
            _release_host_buffers(hb)
            self.last_tile_stats = torch.stack([st[3].view(n_wf, per, -1) for st in state], dim=1).reshape(n0, -1)
        return images, labels, order

    def _finish(self, images, labels, order, patch_size, stretch, flag_sigma, normalize_before_stretch,
                normalize_after_stretch, augmentation_rotations, patch_src=None):
        self._patch_src = patch_src
        self.order = order  # canonical index of every output patch (not in the reference)
        self.patch_flags = labels
        self._patches = None  # rebuilt on first access (see the `patches` property)
        metadata = {  # :394-402
            "patch_size": patch_size,
            "stretch": stretch,
            "flag_sigma": flag_sigma,
            "normalize_before_stretch": normalize_before_stretch,
            "normalize_after_stretch": normalize_after_stretch,
            "augmentation_rotations": augmentation_rotations,
            "original_shapes": getattr(self, "original_shapes", None),
        }
        self.dataset = TorchDataset(images, labels, metadata)
        logger.info("Dataset ready: %d samples", len(self.dataset))
        return self.dataset


def iter_dataset_chunks(data, flags=None, *, chunk_baselines, magnitude=False, device=None, pin=True,
                        lookahead=1, **create_kw):
    """Stream a cube whose output does not fit in HBM at once (BASELINE config 5: 44 baselines x 4
    pols x 1024 x 16384 per GPU give 153 GB of patches) through `create_dataset` in contiguous
    baseline chunks.  Yields `(b0, b1, dataset)`; release `dataset` (or hand it to `BatchWriter`)
    before taking the next chunk and the caching allocator reuses the same output blocks -- the
    "output ring" of SURVEY.md section 8d.

    Every chunk is one `Preprocessor` over `data[b0:b1]` -- the reference's own unit of
    parallelism (one Preprocessor per sample per worker, synthetic_generator.py:55-107) and the
    same semantics as the per-rank baseline shards (`utils.sharding.baseline_shard`): blank-patch
    removal and the shuffle are per chunk, and each chunk draws one `np.random.permutation` from
    the global legacy generator, in chunk order.  Host input is uploaded chunk by chunk (pinned,
    overlapped with phase 1), so the full cube never has to be resident either.

    `lookahead` chunks are kept in flight through `create_dataset_async`: phase 1 (and the upload)
    of chunk c + 1 .. c + lookahead is enqueued before chunk c's host phase, which therefore costs
    no GPU time.  `lookahead=0` is the plain sequential loop.  The permutations are still drawn in
    chunk order (they are drawn in `result()`)."""
    from collections import deque

    n_bl = data.shape[0] if data.ndim == 4 else 1
    if data.ndim == 3:
        data = data[None, ...]
    if chunk_baselines < 1:
        raise ValueError("chunk_baselines must be >= 1")
    if lookahead < 0:
        raise ValueError("lookahead must be >= 0")
    pending = deque()
    for b0 in range(0, n_bl, int(chunk_baselines)):
        b1 = min(n_bl, b0 + int(chunk_baselines))
        pre = Preprocessor(data[b0:b1], None if flags is None else flags[b0:b1],
                           magnitude=magnitude, device=device, pin=pin)
        pending.append((b0, b1, pre.create_dataset_async(**create_kw)))
        if len(pending) > lookahead:
            c0, c1, pd = pending.popleft()
            yield c0, c1, pd.result()
    while pending:
        c0, c1, pd = pending.popleft()
        yield c0, c1, pd.result()
