"""GPUPreprocessor -- mirror of rfi_toolbox/preprocessing/preprocessor.py:784-972: raw complex
patches (no channel extraction, no normalisation, no augmentation), blank removal and shuffle.

Same constructor checks (`ValueError` for ndim not in {3, 4} and for real input, :825-836), same
arguments and the same draws from the GLOBAL legacy generator (`np.random.choice` when
`num_patches` truncates, then one `np.random.permutation`, :918-928).  The upstream method can only
run its "whole waterfall" branch (:885-890): `_create_patches` returns one list (:972) where two
values are unpacked (:893, :896).  Here the patchifying branch does what that code evidently
intends -- non-overlapping P x P tiles, data and flags tiled alike -- with the reference's two
treatments of dimensions that are not multiples of P (:954-970): `num_workers > 0` (the default) is
its worker-pool route through `_patchify_single_waterfall` (:46-112), which zero-pads bottom / right
(the padded samples are unflagged), `num_workers=0` its in-process `patchify` loop, which drops the
remainders (logged).

B200 differences: the tiles are counted and gathered on the device (`rfi_raw_tile_counts`,
`rfi_raw_gather`; each kept tile written once at its final shuffled position); the result is a
pair of device tensors `(N, H, W)` complex and `(N, H, W)` bool -- iterating them yields the
per-patch arrays the reference returns as lists.  `num_workers` only selects pad / drop (above).
"""
from __future__ import annotations

import ctypes as C
import logging

import numpy as np
import torch

from .. import _native
from ..utils.device import as_device_tensor, current_stream_ptr, require_cuda

logger = logging.getLogger(__name__)


class GPUPreprocessor:
    def __init__(self, data, flags=None, *, device=None):
        if data.ndim == 4:
            self.data = data
        elif data.ndim == 3:
            self.data = data[None, ...]
        else:
            raise ValueError(f"Data must be 3D or 4D, got shape {tuple(data.shape)}")
        is_complex = data.is_complex() if isinstance(data, torch.Tensor) else np.iscomplexobj(data)
        if not is_complex:
            raise ValueError("GPUPreprocessor requires complex data. "
                             "Use standard Preprocessor for real-valued data.")
        self.flags = flags
        self.raw_patches = None
        self.raw_masks = None
        self._device = device

    def create_raw_patches(self, patch_size=256, remove_blank=True, num_patches=None, num_workers=4):
        lib = _native.load()
        device = None
        for x in (self.data, self.flags):
            if isinstance(x, torch.Tensor) and x.is_cuda:
                device = x.device
        device = require_cuda(device if device is not None else self._device)
        data = as_device_tensor(self.data, device)
        if data.dtype not in (torch.complex64, torch.complex128):
            raise TypeError(f"unsupported data dtype {data.dtype}: complex64 / complex128")
        dtype = _native.RFI_C64 if data.dtype == torch.complex64 else _native.RFI_C128
        B, npol, C_, T_ = data.shape
        flags = None
        if self.flags is not None:
            flags = as_device_tensor(self.flags, device)
            if flags.ndim == 3:
                flags = flags[None, ...]
            if tuple(flags.shape) != tuple(data.shape):
                raise ValueError(f"flags shape {tuple(flags.shape)} != data shape {tuple(data.shape)}")
            flags = (flags != 0).view(torch.uint8) if flags.dtype != torch.bool else flags.view(torch.uint8)
        P_ = int(patch_size)
        shape_in = (C_, T_)
        if not (C_ <= P_ and T_ <= P_) and (C_ % P_ or T_ % P_):
            if num_workers and num_workers > 0:   # :80-101: zero pad bottom / right (a short dimension up to P)
                pr = (-C_) % P_ if C_ >= P_ else P_ - C_
                pc = (-T_) % P_ if T_ >= P_ else P_ - T_
                # (padded through the (re, im) view: constant padding of complex tensors is not implemented everywhere)
                data = torch.view_as_complex(torch.nn.functional.pad(torch.view_as_real(data), (0, 0, 0, pc, 0, pr)))
                if flags is not None:
                    flags = torch.nn.functional.pad(flags, (0, pc, 0, pr))
                C_, T_ = C_ + pr, T_ + pc
            else:
                logger.info("[GPUPreprocessor] num_workers=0: %d of %d samples per waterfall fall outside the "
                            "%dx%d tiling and are dropped", C_ * T_ - (C_ // P_) * (T_ // P_) * P_ * P_, C_ * T_, P_, P_)
        rows, cols = C.c_int32(), C.c_int32()
        n_tiles = int(lib.rfi_raw_num_tiles(dtype, B * npol, C_, T_, int(patch_size), C.byref(rows), C.byref(cols)))
        if n_tiles < 0:
            _native.check(_native.RFI_E_INVALID, "rfi_raw_num_tiles")
        H, W = rows.value, cols.value
        if not (C_ <= patch_size and T_ <= patch_size):
            self.original_shapes = [shape_in] * (B * npol)
        with torch.cuda.device(device):
            stream = current_stream_ptr(device)
            fptr = flags.data_ptr() if flags is not None else None
            keep = np.ones(n_tiles, dtype=bool)
            if remove_blank and n_tiles:
                counts = torch.empty(n_tiles, dtype=torch.int32, device=device)
                _native.check(lib.rfi_raw_tile_counts(data.data_ptr(), dtype, fptr, B * npol, C_, T_, int(patch_size),
                                                      counts.data_ptr(), stream), "rfi_raw_tile_counts")
                keep = counts.cpu().numpy() > 0
            kept = np.flatnonzero(keep)                         # canonical order of the survivors
            if num_patches and num_patches < len(kept):         # :918-922
                kept = kept[np.random.choice(len(kept), num_patches, replace=False)]
            kept = kept[np.random.permutation(len(kept))]       # :925-928
            dest = np.full(max(n_tiles, 1), -1, dtype=np.int64)
            dest[kept] = np.arange(len(kept))
            dest_dev = torch.from_numpy(dest).to(device)
            patches = torch.empty((len(kept), H, W), dtype=data.dtype, device=device)
            masks = torch.empty((len(kept), H, W), dtype=torch.uint8, device=device)
            if n_tiles:
                _native.check(lib.rfi_raw_gather(data.data_ptr(), dtype, fptr, B * npol, C_, T_, int(patch_size),
                                                 dest_dev.data_ptr(), patches.data_ptr(), masks.data_ptr(), stream),
                              "rfi_raw_gather")
        self.order = kept  # canonical tile index of every output patch (not in the reference)
        self.raw_patches, self.raw_masks = patches, masks.view(torch.bool)
        logger.info("[GPUPreprocessor] %d raw patches of %dx%d", len(kept), H, W)
        return self.raw_patches, self.raw_masks

    def _estimate_storage_mb(self):
        if self.raw_patches is None or len(self.raw_patches) == 0:
            return 0
        return self.raw_patches.element_size() * self.raw_patches.numel() / (1024 * 1024)
