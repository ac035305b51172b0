set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --workload pyramidal_4k --batch 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b4_v3.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b4_v3.log | cut -c1-200
python bench.py --workload pyramidal_4k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b16_v3.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b16_v3.log | cut -c1-200
python bench.py --workload pyramidal_8k --batch 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8k_v3.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_8k_v3.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pyr3.csv python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyr3.log 2>&1; echo ncu rc=$?
