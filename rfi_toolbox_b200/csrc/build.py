"""Build librfi_b200.so in-tree with nvcc for sm_100a (no torch involved: the library is a
plain C-ABI shared object, see include/rfi_b200.h).

    python -m rfi_toolbox_b200.csrc.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
LIB = PKG / "_lib" / "librfi_b200.so"
OBJ = HERE / "build"
SOURCES = ["rfi_error.cu", "rfi_tiles.cu", "rfi_generic.cu", "rfi_bigtile.cu", "rfi_metrics.cu", "rfi_stats.cu", "rfi_gstats.cu", "rfi_pairs.cu", "rfi_synth.cu", "rfi_raw.cu", "rfi_ingest.cu", "rfi_host.cpp"]
HEADERS = sorted(HERE.glob("*.cuh")) + [PKG.parent / "include" / "rfi_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # one IEEE op per source op: parity with NumPy (DESIGN.md numerics)
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: librfi_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [HERE / s for s in SOURCES if (HERE / s).exists()]
    if not force and not _stale(LIB, srcs + HEADERS + [Path(__file__)]):
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src] + HEADERS + [Path(__file__)]):
            if src.suffix == ".cpp":  # host-only helper: plain g++ (vectorises what nvcc's host pass does not)
                cmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-c", str(src), "-o", str(obj)]
            else:
                cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"compilation failed for {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
