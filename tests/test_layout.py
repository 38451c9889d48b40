"""Repository rules: the product never imports the oracle, nothing run on the GPU box reads
/root/reference, and no forbidden CUDA batch-copy call is named anywhere."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _py(dirname):
    return [p for p in (ROOT / dirname).rglob("*.py")]


def test_product_does_not_import_oracle():
    for p in _py("rfi_toolbox_b200"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_only_allowed_files_import_oracle():
    allowed = {"bench.py", "__graft_entry__.py"}
    for p in ROOT.glob("*.py"):
        if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), flags=re.M):
            assert p.name in allowed, p


def test_gpu_side_never_reads_reference():
    names = ["bench.py", "__graft_entry__.py", "tests/test_gpu_parity.py", "tests/test_gpu_metrics.py",
             "tests/cubes.py", "tests/golden_util.py"]
    for n in names:
        assert "/root/reference" not in (ROOT / n).read_text(), n
    for p in _py("rfi_toolbox_b200") + _py("oracle"):
        assert "sys.path" not in p.read_text() or "/root/reference" not in p.read_text(), p


def test_no_batched_memcpy_calls():
    bad = re.compile(r"cu(da)?Memcpy(3D)?BatchAsync")
    for p in list(ROOT.rglob("*.cu")) + list(ROOT.rglob("*.cuh")) + list(ROOT.rglob("*.py")) + list(ROOT.rglob("*.h")):
        if "gpurun_out" in p.parts or p.name == "test_layout.py":
            continue
        assert not bad.search(p.read_text(errors="ignore")), p
