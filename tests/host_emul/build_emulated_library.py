#!/usr/bin/env python3
"""Build tests/host_emul/_build/libof_b200_emulated.so: every product .cu compiled by g++ on top of cuda_on_host.h,
linked with fake_cudart.cpp -- the C ABI of include/of_b200.h with the CPU standing in for the device.

TEST INFRASTRUCTURE ONLY.  The library is never placed next to the product's libof_b200.so; tests load it explicitly
(OF_B200_LIB_NAME, the wrapper's hook for experimental builds) in a child process.

    python tests/host_emul/build_emulated_library.py [--force]
"""
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "optical-flow-fpga_b200" / "csrc"
CUDA_INC = Path("/usr/local/cuda/include")
OUT = HERE / "_build" / "libof_b200_emulated.so"
UNITS = ["emul_of_api", "emul_lk_march", "emul_lk_tile", "emul_lk_tile5", "emul_lk_exact_march", "emul_pyramid", "emul_pyramid_march", "emul_motion",
         "emul_fixed", "emul_metrics", "emul_peer", "fake_cudart"]


def build(force: bool = False) -> Path:
    gxx = shutil.which("g++")
    if gxx is None or not (CUDA_INC / "cuda_runtime.h").exists():
        raise RuntimeError("needs g++ and the CUDA headers")
    deps = [HERE / f"{u}.cpp" for u in UNITS] + [HERE / "cuda_on_host.h"] + sorted(CSRC.iterdir()) + [ROOT / "include" / "of_b200.h"]
    if not force and OUT.exists() and all(d.stat().st_mtime <= OUT.stat().st_mtime for d in deps):
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    objs = []
    procs = []
    for u in UNITS:
        obj = OUT.parent / f"{u}.o"
        cmd = [gxx, "-O1", "-ffp-contract=off", "-frounding-math", "-std=c++17", "-fPIC", "-pthread", "-w", "-DOF_HOST_EMULATION",
               "-DCUDA_ON_HOST_FAKE_RUNTIME", "-DEMUL_EXACT_MARCH_NO_TILE_GEOMETRY", "-I", str(CSRC), "-I", str(CUDA_INC), "-c", str(HERE / f"{u}.cpp"), "-o", str(obj)]
        procs.append((u, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
        objs.append(str(obj))
    for u, p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"{u}: {err[-3000:]}")
    res = subprocess.run([gxx, "-shared", "-pthread", "-o", str(OUT), *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stderr[-3000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
