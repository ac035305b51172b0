#!/bin/bash
# Round 2, last check at head (one GPU): suite, smoke, the default bench line
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.log 2>&1; echo rc=$?; grep '^{' gpurun_out/bench_default.log | cut -c1-300
