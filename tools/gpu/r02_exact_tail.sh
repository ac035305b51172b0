#!/bin/bash
# Round 2: exact marching kernel with the fused iteration tail -- GPU suite + the exact workloads (one GPU)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in pyramidal_4k_exact pyramidal_8k_exact; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
