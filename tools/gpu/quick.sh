#!/bin/bash
# a short A/B round trip: GPU tests, the pyramidal bench line and its launch list
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py --workload pyramidal_4k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyramidal_4k.log 2>&1; tail -1 gpurun_out/bench_pyramidal_4k.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pyr.csv python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr.log 2>&1; echo ncu rc=$?
grep -E "warp_rows|pyramid_march|upsample" gpurun_out/launches_pyr.csv | tail -12 | awk -F'","' '{print substr($5,1,40), $9, $15}'
