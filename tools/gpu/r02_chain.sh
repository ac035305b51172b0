#!/bin/bash
# Round 2: refinement kernel that warps the next iteration's input itself (one GPU):  gpurun --timeout 1500 -- 'bash tools/gpu/r02_chain.sh'
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in pyramidal_4k pyramidal_8k pyramidal_4k_w7; do
  OF_B200_REFINE_WARP=chain timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
  OF_B200_REFINE_WARP=rows timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_rows.log 2>&1; echo "$wl (warp_rows per iteration) rc=$?"; grep '^{' gpurun_out/bench_${wl}_rows.log | cut -c1-200
done
OF_B200_REFINE_WARP=chain timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|warp_rows|pyramid|upsample" -c 200 --csv --log-file gpurun_out/launches_pyr4k_b4_chain.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_chain.log 2>&1; echo ncu rc=$?
OF_B200_REFINE_WARP=chain timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lk_march_kernel" --launch-skip 28 --launch-count 1 -o gpurun_out/prof_refine_chain -f python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_refine_chain.log 2>&1; echo ncu rc=$?
