"""NVLink peer-memory exchange buffers for the fused count + all-reduce kernel
(`rfi_confusion_counts_allreduce`, csrc/rfi_metrics.cu).

One process per GPU (torch.distributed).  Every rank allocates a small exchange buffer through
the C ABI (`rfi_peer_alloc`), the 64-byte CUDA IPC handles travel over the process group
(`all_gather_object` -- plumbing), and every rank opens its peers' buffers (`rfi_peer_open`).
After that a metrics call is ONE kernel on the caller's stream: no NCCL launch.  Ranks on
different hosts, or without CUDA IPC / peer access, keep the NCCL all-reduce (logged once).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import socket

import torch

from .. import _native

logger = logging.getLogger(__name__)
_CONTEXTS = {}
MAX_WORLD = 16


class PeerContext:
    """Exchange buffers of one (group, device).  Construction is COLLECTIVE and cannot diverge:
    every rank always takes part in the one `all_gather_object` (publishing a failure marker
    instead of raising before it) and in one `all_reduce(MIN)` of a success flag after the open
    step, so either every rank ends up with `ok = True` or every rank ends up with `ok = False`
    (and the caller all-reduces over NCCL on all of them).  On failure everything this rank
    allocated or opened is released again."""

    def __init__(self, group, device):
        import torch.distributed as dist
        lib = _native.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.epoch = 0
        self.ok = False
        self.reason = None
        self.ptrs = None
        self._own = C.c_void_p()
        self._opened = []
        with torch.cuda.device(device):
            handle, err = (C.c_ubyte * 64)(), None
            try:
                if self.world > MAX_WORLD:
                    raise RuntimeError(f"peer exchange supports at most {MAX_WORLD} ranks")
                if os.environ.get("RFI_PEER_FAIL_RANK") == str(self.rank):  # fault injection (tests)
                    raise RuntimeError("injected failure (RFI_PEER_FAIL_RANK)")
                _native.check(lib.rfi_peer_alloc(C.byref(self._own), handle), "rfi_peer_alloc")
            except Exception as exc:  # published below: the other ranks must not wait for us
                err = f"rank {self.rank}: {exc}"
            mine = (socket.gethostname(), bytes(handle) if err is None else None, err)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            errs = [e for _, _, e in everyone if e]
            if not errs and len({h for h, _, _ in everyone}) != 1:
                errs = ["ranks span several hosts"]
            ptrs = []
            if not errs:
                try:
                    for r, (_, h, _) in enumerate(everyone):
                        if r == self.rank:
                            ptrs.append(self._own.value)
                            continue
                        p = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(h)
                        _native.check(lib.rfi_peer_open(buf, C.byref(p)), "rfi_peer_open")
                        self._opened.append(p)
                        ptrs.append(p.value)
                except Exception as exc:
                    errs = [f"rank {self.rank}: {exc}"]
            # agreement: the peer path is used only if EVERY rank mapped every buffer
            flag = torch.tensor([0 if errs else 1], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 1:
                self.ptrs = (C.c_void_p * self.world)(*ptrs)
                self.ok = True
            else:
                self.reason = "; ".join(errs) if errs else "a peer rank could not map the exchange buffers"
                self.close()
        dist.barrier(group=group)  # every buffer is mapped everywhere before its first use

    def close(self):
        """Unmap the peers' buffers and free this rank's own (idempotent)."""
        lib = _native.load()
        for p in self._opened:
            try:
                lib.rfi_peer_close(p)
            except Exception:
                pass
        self._opened = []
        if self._own.value:
            try:
                lib.rfi_peer_free(self._own)
            except Exception:
                pass
            self._own = C.c_void_p()

    def next_epoch(self):
        self.epoch += 1
        return self.epoch


def peer_context(group, device):
    """The exchange context of (`group`, `device`), built on first use (collectively: every rank of
    the group must call this); None if the ranks cannot share memory -- on ALL ranks alike -- and
    the caller then all-reduces over NCCL."""
    key = (id(group) if group is not None else 0, device.index)
    if key not in _CONTEXTS:
        ctx = PeerContext(group, device)
        if not ctx.ok:
            logger.warning("peer-memory exchange unavailable (%s): metric counts go through NCCL", ctx.reason)
            ctx = None
        _CONTEXTS[key] = ctx
    return _CONTEXTS[key]
