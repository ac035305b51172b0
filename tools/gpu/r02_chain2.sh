#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pyramidal or rowband or large_window" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in pyramidal_4k pyramidal_8k; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|warp_rows" -c 40 --csv --log-file gpurun_out/launches_pyr4k_b4_chain2.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_chain2.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_march_kernel" --launch-skip 6 --launch-count 1 -o gpurun_out/prof_refine_chain2 -f python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_refine_chain2.log 2>&1; echo ncu rc=$?
