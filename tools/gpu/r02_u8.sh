#!/bin/bash
# Round 2: uint8 / fixed-point flavours of the marching kernel after a change (one GPU)
set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "u8 or uint8 or fixed or fx or smoke or verifier_pipeline" > gpurun_out/pytest_gpu_u8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_u8.log
for wl in single_1080p_u8 fixed_1080p; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/prof_march_u8_r02 -f python bench.py --workload single_1080p_u8 --workloads none --steps 4 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march_u8.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/prof_march_fx_r02 -f python bench.py --workload fixed_1080p --workloads none --steps 4 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march_fx.log 2>&1; echo ncu rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|lk_tile|lk_exact|pyramid|warp_rows|upsample|select_copy|iter_finalize" -c 80 --csv --log-file gpurun_out/launches_single.csv python bench.py --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_single.log 2>&1; echo ncu rc=$?
