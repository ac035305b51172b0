#!/bin/bash
# The kernels' sources run on the CPU (tests/test_kernel_host_emulation.py) once more under AddressSanitizer, with the
# dynamic shared memory allocated at exactly the size the launcher asked for: any access past a kernel's shared memory
# or a frame / flow buffer is reported.  TEST INFRASTRUCTURE; compute-sanitizer's stand-in where it is not available.
#   bash tests/host_emul/asan_check.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/../.."
ASAN=$(g++ -print-file-name=libasan.so)
export OF_EMUL_EXTRA_FLAGS="-fsanitize=address -fno-omit-frame-pointer -g -DOF_EMUL_EXACT_SMEM"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1
LD_PRELOAD=$ASAN python -m pytest tests/test_kernel_host_emulation.py -x -q -p no:cacheprovider -k "${1:-march or refine or exact or warp}"
