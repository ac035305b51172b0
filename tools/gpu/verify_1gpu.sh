#!/bin/bash
# What the round-end driver does on one B200, plus the launch lists kept under profiles/:
#   gpurun --timeout 1200 -- 'bash tools/gpu/verify_1gpu.sh'
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_reference.log | cut -c1-250
python bench.py > gpurun_out/bench_default.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_default.log | cut -c1-250
for wl in pyramidal_4k pyramidal_8k single_4k single_1080p_u8 fixed_1080p; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; tail -1 gpurun_out/bench_$wl.log | cut -c1-200
done
# launch lists: the kernels' share of a step (never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_single.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_single.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pyr.csv python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr.log 2>&1; echo ncu rc=$?
