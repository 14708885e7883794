"""Builds libdxt_lossless_transform_cuda.so in-tree with nvcc for sm_100a (B200).

The library is the product: hand-written CUDA kernels plus the C ABI in csrc/cabi.cu.  It is built
next to this file so that it travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libdxt_lossless_transform_cuda.so"


def _lines(name: str) -> list[str]:
    return [l.strip() for l in (CSRC / name).read_text().splitlines() if l.strip() and not l.startswith("#")]


# One list of sources and flags for this build and for the Rust crate's build.rs (rust/dxt-lossless-transform-cuda).
SOURCES = _lines("SOURCES.txt")
NVCC_FLAGS = _lines("NVCC_FLAGS.txt")  # -diag-suppress 177: unused static members of the layout helper in some instantiations


def find_nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")
    return cand


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    built = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.txt")) + [Path(__file__)]
    return any(p.stat().st_mtime > built for p in deps)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    cmd = [nvcc, *NVCC_FLAGS, "-shared", "-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    if verbose:
        print(proc.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
