"""Host-side plumbing shared by the operator mirrors: device placement and dtype codes."""
from .device import as_device_tensor, current_stream_ptr, require_cuda  # noqa: F401
