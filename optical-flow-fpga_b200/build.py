#!/usr/bin/env python3
"""Build libof_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python optical-flow-fpga_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / os.environ.get("OF_B200_LIB_NAME", "libof_b200.so")  # variants: experiments only
SOURCES = ["of_api.cu", "lk_march.cu", "lk_tile.cu", "lk_tile5.cu", "pyramid.cu", "pyramid_march.cu", "peer.cu", "metrics.cu", "motion.cu", "lk_fixed.cu"]
HEADERS = ["of_common.cuh", "of_kernels.h", "of_rowband.inl", "warp_rows.cuh", "../../include/of_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # reference arithmetic has no FMA contraction; fast kernels call fmaf() explicitly
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not LIB.exists():
        return False
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / h for h in HEADERS] + [Path(__file__)]
    return all(d.stat().st_mtime <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return LIB
    cmd = [find_nvcc(), *NVCC_FLAGS, *os.environ.get("OF_NVCC_DEFS", "").split()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libof_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
