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
SOURCES = ["of_api.cu", "lk_march.cu", "lk_tile.cu", "lk_tile5.cu", "lk_exact_march.cu", "pyramid.cu", "pyramid_march.cu", "peer.cu", "metrics.cu", "motion.cu", "lk_fixed.cu"]
HEADERS = ["of_common.cuh", "f32x2.cuh", "of_kernels.h", "of_rowband.inl", "warp_rows.cuh", "../../include/of_b200.h"]

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


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def up_to_date() -> bool:
    deps = [CSRC / s for s in SOURCES] + [CSRC / h for h in HEADERS] + [Path(__file__)]
    return not _stale(LIB, deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    """One object per source (compiled in parallel, rebuilt only when the source or a header changed), then
    the shared library.  OF_NVCC_DEFS (experiments) changes the flags, so it forces a full rebuild."""
    if not force and up_to_date() and not os.environ.get("OF_NVCC_DEFS"):
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = find_nvcc()
    extra = os.environ.get("OF_NVCC_DEFS", "").split()
    objdir = HERE / "build" / (LIB.stem + ("_defs" if extra else ""))
    objdir.mkdir(parents=True, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])
    headers = [CSRC / h for h in HEADERS] + [Path(__file__)]

    def compile_one(src: str):
        obj = objdir / (Path(src).stem + ".o")
        if not (force or extra) and not _stale(obj, [CSRC / src] + headers):
            return obj, None
        res = subprocess.run([nvcc, *flags, "-c", "-o", str(obj), str(CSRC / src)], capture_output=True, text=True)
        return obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    failed = False
    for _, res in results:
        if res is not None and (verbose or res.returncode != 0):
            sys.stderr.write(res.stdout + res.stderr)
        failed |= res is not None and res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libof_b200.so")
    link = [nvcc, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)]
    res = subprocess.run(link + [str(o) for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libof_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
