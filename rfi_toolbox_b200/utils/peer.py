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
import socket

import torch

from .. import _native

logger = logging.getLogger(__name__)
_CONTEXTS = {}
MAX_WORLD = 16


class PeerContext:
    def __init__(self, group, device):
        import torch.distributed as dist
        lib = _native.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.epoch = 0
        self._own = C.c_void_p()
        self._opened = []
        if self.world > MAX_WORLD:
            raise RuntimeError(f"peer exchange supports at most {MAX_WORLD} ranks")
        with torch.cuda.device(device):
            handle = (C.c_ubyte * 64)()
            _native.check(lib.rfi_peer_alloc(C.byref(self._own), handle), "rfi_peer_alloc")
            mine = (socket.gethostname(), bytes(handle))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            if len({h for h, _ in everyone}) != 1:
                raise RuntimeError("ranks span several hosts")
            ptrs = []
            for r, (_, h) in enumerate(everyone):
                if r == self.rank:
                    ptrs.append(self._own.value)
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _native.check(lib.rfi_peer_open(buf, C.byref(p)), "rfi_peer_open")
                self._opened.append(p)
                ptrs.append(p.value)
            self.ptrs = (C.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)  # every buffer is mapped everywhere before its first use

    def next_epoch(self):
        self.epoch += 1
        return self.epoch


def peer_context(group, device):
    """The exchange context of (`group`, `device`), built on first use; None if the ranks cannot
    share memory (then the caller all-reduces over NCCL)."""
    key = (id(group) if group is not None else 0, device.index)
    if key not in _CONTEXTS:
        try:
            _CONTEXTS[key] = PeerContext(group, device)
        except Exception as exc:  # CUDA IPC unavailable, several hosts, ...
            logger.warning("peer-memory exchange unavailable (%s): metric counts go through NCCL", exc)
            _CONTEXTS[key] = None
    return _CONTEXTS[key]
