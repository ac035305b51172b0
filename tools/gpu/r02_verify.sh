#!/bin/bash
# Round 2, one B200: the GPU suite, smoke, the default bench line (with the `workloads` map) and the reference arm.
#   gpurun --timeout 1500 -- 'bash tools/gpu/r02_verify.sh'
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu.txt 2>&1; nproc >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.log 2>&1; echo rc=$?; tail -5 gpurun_out/bench_default.log | cut -c1-400
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.log 2>&1; echo rc=$?; tail -5 gpurun_out/bench_reference.log | cut -c1-300
