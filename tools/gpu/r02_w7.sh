#!/bin/bash
# Round 2: window-7 marching kernels + warp_rows address trim (one GPU):  gpurun --timeout 1200 -- 'bash tools/gpu/r02_w7.sh'
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in single_1080p_w7 pyramidal_4k_w7 pyramidal_4k pyramidal_8k; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
# the same two window-7 workloads on the kernels they used before (first tile kernel): OF_B200_MARCH7=off
for wl in single_1080p_w7 pyramidal_4k_w7; do
  OF_B200_MARCH7=off timeout 300 python bench.py --workload $wl --workloads none --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_tile.log 2>&1; echo "$wl (tile) rc=$?"; grep '^{' gpurun_out/bench_${wl}_tile.log | cut -c1-200
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|warp_rows|pyramid|upsample" -c 200 --csv --log-file gpurun_out/launches_pyr4k_w7_b4.csv python bench.py --workload pyramidal_4k_w7 --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_w7.log 2>&1; echo ncu rc=$?
