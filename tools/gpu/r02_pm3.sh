#!/bin/bash
# Round 2: classic refinement with a 2-stage ring; pyramid kernel at 3 CTAs/SM (variant library) -- one GPU
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in pyramidal_4k pyramidal_8k pyramidal_4k_exact; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
  OF_B200_LIB_NAME=libof_b200_pm3.so timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_pm3.log 2>&1; echo "$wl (pyramid kernel at 3 CTAs/SM) rc=$?"; grep '^{' gpurun_out/bench_${wl}_pm3.log | cut -c1-200
done
OF_B200_LIB_NAME=libof_b200_pm3.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pyramid" -c 12 --csv --log-file gpurun_out/launches_pyr4k_b4_pm3.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pm3.log 2>&1; echo ncu rc=$?
