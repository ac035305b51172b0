set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload pyramidal_4k --batch 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b4_march.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b4_march.log | cut -c1-400
OF_B200_PYRAMID=tile python bench.py --workload pyramidal_4k --batch 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b4_tile.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b4_tile.log | cut -c1-200
python bench.py --workload pyramidal_4k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b16_march.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b16_march.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_pyr2.csv python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr2.log 2>&1; echo ncu rc=$?
