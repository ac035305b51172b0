#!/bin/bash
# Round 2: pyramidal workloads after a kernel change (one GPU):  gpurun --timeout 900 -- 'bash tools/gpu/r02_pyr.sh'
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in pyramidal_4k pyramidal_4k_exact pyramidal_8k; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pyr4k_b4.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr.log 2>&1; echo ncu rc=$?
