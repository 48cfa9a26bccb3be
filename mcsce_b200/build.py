"""Build the in-tree CUDA library (nvcc, sm_100a) and the oracle's C restatement (gcc).

    python -m mcsce_b200.build

``libmcsce_b200.so`` is written next to the package so it travels with the tree; it is
never installed into site-packages.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CU_SOURCES = [PKG / "csrc" / "mcsce_kernels.cu", PKG / "csrc" / "hpmf_kernels.cu", PKG / "csrc" / "pdb_writer.cpp"]
LIB = PKG / "libmcsce_b200.so"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


def _stale(target, sources):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def build_cuda(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    deps = CU_SOURCES + [ROOT / "include" / "mcsce_b200.h"] + sorted((PKG / "csrc").glob("*.cuh"))    # (headers included by the sources)
    if not force and not _stale(LIB, deps):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + [str(s) for s in CU_SOURCES]
    subprocess.run(cmd, check=True)
    return LIB


def build_oracle_c(force=False):
    """gcc build of oracle/growth_c (test infrastructure; building it is not using it)."""
    src = ROOT / "oracle" / "growth_c.c"
    out = ROOT / "oracle" / "libgrowth_c.so"
    if not src.exists():
        return None
    if force or _stale(out, [src]):
        subprocess.run(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC", "-fopenmp",
                        "-o", str(out), str(src), "-lm"], check=True)
    return out


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_oracle_c(force="--force" in sys.argv))
