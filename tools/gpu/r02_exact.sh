#!/bin/bash
# Round 2: the exact marching kernel on the GPU -- suite, A/B bench against the tile kernel, ncu capture.
#   gpurun --timeout 1200 -- 'bash tools/gpu/r02_exact.sh'
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for wl in single_1080p_exact pyramidal_4k_exact; do
  OF_B200_EXACT=march timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_march.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_${wl}_march.log | cut -c1-160
  OF_B200_EXACT=tile timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_tile.log 2>&1; echo "$wl tile rc=$?"; grep '^{' gpurun_out/bench_${wl}_tile.log | cut -c1-160
done
OF_B200_EXACT=march timeout 300 ncu --set full --clock-control none --import-source on -k regex:lk_exact_march --launch-skip 3 --launch-count 1 -o gpurun_out/prof_exact_march -f python bench.py --workload single_1080p_exact --workloads none --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_exact_march.log 2>&1; echo ncu rc=$?
