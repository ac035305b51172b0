#!/bin/bash
# Round 2, head of the branch on one B200: what the round-end driver does (suite, smoke, both bench arms), then the
# launch list and the full capture of the dominant kernel of the same bench command (each only after the command
# itself exited 0 without ncu).   gpurun --timeout 1500 -- 'bash tools/gpu/r02_final.sh'
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/bench_reference.log 2>&1; echo rc=$?; grep '^{' gpurun_out/bench_reference.log | cut -c1-200
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default.log 2>&1; echo rc=$?; grep '^{' gpurun_out/bench_default.log | cut -c1-300
timeout 300 python tools/device_verifier.py > gpurun_out/device_verifier.txt 2>&1; echo "device_verifier rc=$?"; tail -3 gpurun_out/device_verifier.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|lk_tile|lk_exact|pyramid|warp_rows|upsample|select_copy|iter_finalize" -c 80 --csv --log-file gpurun_out/launches_single.csv python bench.py --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_single.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/prof_march_r02 -f python bench.py --workloads none --steps 4 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march.log 2>&1; echo ncu rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lk_march|lk_tile|lk_exact|pyramid|warp_rows|upsample|select_copy|iter_finalize" -c 900 --csv --log-file gpurun_out/launches_pyr8k_b1.csv python bench.py --workload pyramidal_8k --batch 1 --workloads none --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr8k.log 2>&1; echo ncu rc=$?
