#!/bin/bash
# Round 2: GPU suite + bench lines of the workloads touched by the kernel changes + ncu captures (one GPU).
#   gpurun --timeout 1500 -- 'bash tools/gpu/r02_kernels.sh'
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for wl in single_1080p_exact pyramidal_4k pyramidal_4k_exact pyramidal_8k; do
  timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
# launch list of the fast pyramidal step (share of the step per kernel; never a bench value)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pyr4k_b4.csv python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr.log 2>&1; echo ncu rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_pyr4k_exact_b4.csv python bench.py --workload pyramidal_4k_exact --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr_exact.log 2>&1; echo ncu rc=$?
# full captures: the exact tile kernel, the float32 pyramid kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_tile5 --launch-skip 3 --launch-count 1 -o gpurun_out/prof_tile5_v3 -f python bench.py --workload single_1080p_exact --workloads none --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tile5.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pyramid_march --launch-skip 16 --launch-count 1 -o gpurun_out/prof_pyrmarch_f32 -f python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyrmarch.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:warp_rows --launch-skip 30 --launch-count 1 -o gpurun_out/prof_warprows_f32 -f python bench.py --workload pyramidal_4k --batch 4 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_warprows.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/*.ncu-rep | tail -5
