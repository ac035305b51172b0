set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_default.log | cut -c1-250
ncu --set full --clock-control none --import-source on -k regex:pyramid_march --launch-skip 4 --launch-count 1 -o gpurun_out/prof_pyrmarch_v3 -f python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyrmarch_v3.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_march_fx -f python bench.py --workload fixed_1080p --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march_fx.log 2>&1; echo ncu rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_8k.csv python bench.py --workload pyramidal_8k --batch 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_8k.log 2>&1; echo ncu rc=$?
