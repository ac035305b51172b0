#!/bin/bash
# Round 2: launch list of one exact pyramidal step (4 x 4K, 3 x 3) at head -- one GPU
set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|lk_tile|lk_exact|pyramid|warp_rows|upsample|select_copy|iter_finalize" -c 200 --csv --log-file gpurun_out/launches_pyr4k_exact_b4.csv python bench.py --workload pyramidal_4k_exact --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_exact.log 2>&1; echo ncu rc=$?
