#!/bin/bash
# The kernels' sources run on the CPU (tests/test_kernel_host_emulation.py) once more under a sanitizer -- the stand-in
# for compute-sanitizer where that is not available.  TEST INFRASTRUCTURE.
#   bash tests/host_emul/sanitize.sh asan [pytest -k expression]   AddressSanitizer; the dynamic shared memory of every
#        launch is allocated at exactly the size the launcher asked for, the frame / flow buffers are NumPy allocations
#        it tracks: an access past a kernel's shared memory, ring stage or image plane is reported (memcheck)
#   bash tests/host_emul/sanitize.sh tsan [pytest -k expression]   ThreadSanitizer; CUDA threads are OS threads, barriers
#        and mbarriers are pthread barriers / a mutex and atomics: an unordered pair of accesses to shared or global
#        memory -- a missing __syncwarp, a stage handed over too early -- is reported (racecheck)
set -e
cd "$(dirname "$0")/../.."
case "$1" in
  asan) LIB=$(g++ -print-file-name=libasan.so)
        export OF_EMUL_EXTRA_FLAGS="-fsanitize=address -fno-omit-frame-pointer -g -DOF_EMUL_EXACT_SMEM"
        export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:halt_on_error=1 ;;
  tsan) LIB=$(g++ -print-file-name=libtsan.so)
        export OF_EMUL_EXTRA_FLAGS="-fsanitize=thread -g"
        export TSAN_OPTIONS=halt_on_error=0:report_signal_unsafe=0:exitcode=0 ;;
  *) echo "usage: $0 asan|tsan [-k expression]"; exit 2 ;;
esac
LOG=$(mktemp)
LD_PRELOAD=$LIB python -m pytest tests/test_kernel_host_emulation.py -x -q -p no:cacheprovider -k "${2:-march or refine or exact or warp}" 2>&1 | tee "$LOG" | tail -3
N=$(grep -c "WARNING: ThreadSanitizer\|ERROR: AddressSanitizer" "$LOG" || true)
echo "sanitizer reports: $N"
test "$N" = "0"
