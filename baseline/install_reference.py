#!/usr/bin/env python
"""Install the unmodified reference package into baseline/_ref (build container only).

    python baseline/install_reference.py

`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`:
the package is pure Python; `--no-deps` because the offline wheelhouse holds no second copy of
numpy / scipy (both are already importable) and dependency resolution is the only step that
fails without it; from a copy under /tmp because the build writes egg-info into the source
tree, which is read-only.  A no-op when baseline/_ref is already populated or when the
reference tree is absent (the GPU box: the prebuilt baseline/_ref travels with the snapshot)."""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF, DST = Path("/root/reference"), ROOT / "baseline" / "_ref"


def main():
    if (DST / "rfi_toolbox" / "__init__.py").exists() or not REF.exists():
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "reference"
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git"))
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", str(DST), str(src)],
                           capture_output=True, text=True)
    print("reference -> baseline/_ref:", "ok" if r.returncode == 0 else f"FAILED ({r.stderr[-300:]})")
    return r.returncode


if __name__ == "__main__":
    main()
