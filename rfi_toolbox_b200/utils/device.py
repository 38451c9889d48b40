"""Device placement helpers.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import numpy as np
import torch


def require_cuda(device=None) -> torch.device:
    """The product path has no CPU fallback: fail loudly without a CUDA device."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "rfi_toolbox_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback"
        )
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"rfi_toolbox_b200 runs on CUDA devices only (got {device})")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def as_device_tensor(x, device: torch.device, pin: bool = False) -> torch.Tensor:
    """NumPy array / torch tensor (any device) -> contiguous tensor on `device`.
    Host arrays go through one H2D copy (from pinned memory when `pin`)."""
    if isinstance(x, torch.Tensor):
        t = x.detach()
    else:
        a = np.asarray(x)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    if t.device != device:
        if pin and t.device.type == "cpu" and not t.is_pinned():
            t = t.pin_memory()
        t = t.to(device, non_blocking=True)
    return t.contiguous()


def as_host_tensor(x, pin: bool = False):
    """Host array / CPU tensor -> contiguous CPU tensor (pinned when it already is, or when `pin`);
    None for anything already on a device.  Used by the chunked, overlapped upload."""
    if isinstance(x, torch.Tensor):
        if x.device.type != "cpu":
            return None
        t = x.detach().contiguous()
    else:
        a = np.asarray(x)
        if not a.flags.c_contiguous:
            a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
    if pin and not t.is_pinned():
        t = t.pin_memory()
    return t


_COPY_STREAMS = {}


def copy_stream(device: torch.device) -> torch.cuda.Stream:
    """One side stream per device for host->device uploads that overlap with kernels."""
    s = _COPY_STREAMS.get(device.index)
    if s is None:
        s = _COPY_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return s


def bind_host_to_device(device) -> list | None:
    """Pin the calling process to the CPU cores (hence, by first touch, the host memory node)
    closest to `device`, as NVML reports them.  With one process per GPU this keeps every rank's
    pinned upload buffers on its GPU's own socket: 8 concurrent H2D streams otherwise share one
    inter-socket link.  Returns the CPU list, or None when NVML / the topology is unavailable."""
    import os
    try:
        import pynvml
        device = torch.device(device)
        props = torch.cuda.get_device_properties(device)
        bus = f"{props.pci_domain_id:08X}:{props.pci_bus_id:02X}:{props.pci_device_id:02X}.0"
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None
