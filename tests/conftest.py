import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the live reference at /root/reference")


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not REFERENCE.exists():
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load librfi_b200.so."""
    from rfi_toolbox_b200.csrc.build import build
    from rfi_toolbox_b200 import _native

    build()
    return _native.load()


def pytest_sessionfinish(session, exitstatus):
    """Measured image-channel distances (tests/test_gpu_parity.py::image_errors) -> gpurun_out/."""
    try:
        from tests.test_gpu_parity import IMAGE_ERRORS
    except Exception:
        return
    if not IMAGE_ERRORS:
        return
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "image_errors.txt", "w") as fh:
        fh.write("# case | channel | max|numpy-exact64| | max|cuda-exact64| | max|cuda-numpy|\n")
        for label, ch, e_np, e_gpu, e_d in IMAGE_ERRORS:
            fh.write(f"{label or '-'} | {ch} | {e_np:.3e} | {e_gpu:.3e} | {e_d:.3e}\n")
