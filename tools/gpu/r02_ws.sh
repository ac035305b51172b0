#!/bin/bash
# Round 2: warp-specialised refinement kernel (OF_B200_REFINE=ws) against the split form -- one GPU
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_mode_refinement_variants" > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ws.log
for wl in pyramidal_4k pyramidal_8k; do
  OF_B200_REFINE=ws timeout 120 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_ws.log 2>&1; echo "$wl ws rc=$?"; grep '^{' gpurun_out/bench_${wl}_ws.log | cut -c1-200
  timeout 120 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}.log 2>&1; echo "$wl split rc=$?"; grep '^{' gpurun_out/bench_${wl}.log | cut -c1-200
done
OF_B200_REFINE=ws timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|warp_rows" -c 40 --csv --log-file gpurun_out/launches_pyr4k_b4_ws.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_ws.log 2>&1; echo ncu rc=$?
OF_B200_REFINE=ws timeout 200 ncu --set full --clock-control none --import-source on -k regex:"lk_march_kernel" --launch-skip 6 --launch-count 1 -o gpurun_out/prof_refine_ws -f python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_refine_ws.log 2>&1; echo ncu rc=$?
