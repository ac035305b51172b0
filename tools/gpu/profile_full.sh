set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_reference.log | cut -c1-300
python bench.py > gpurun_out/bench_default.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_default.log | cut -c1-300
python bench.py --workload pyramidal_4k --steps 10 --warmup 3 > gpurun_out/bench_pyr4k.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr4k.log | cut -c1-300
python bench.py --workload single_1080p_u8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_single_1080p_u8.log 2>&1; echo rc=$?
# launch lists (kernel share of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_single2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_single2.log 2>&1; echo ncu rc=$?
# full captures
ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_march_v2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march_v2.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:pyramid_march --launch-skip 16 --launch-count 1 -o gpurun_out/prof_pyrmarch_v2 -f python bench.py --workload pyramidal_4k --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyrmarch_v2.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:upsample_flow_tile --launch-skip 7 --launch-count 1 -o gpurun_out/prof_upsample_v2 -f python bench.py --workload pyramidal_4k --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_upsample_v2.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:warp_rows --launch-skip 30 --launch-count 1 -o gpurun_out/prof_warprows_v2 -f python bench.py --workload pyramidal_4k --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_warprows_v2.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 30 --launch-count 1 -o gpurun_out/prof_marchrefine_v2 -f python bench.py --workload pyramidal_4k --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_marchrefine_v2.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:lk_march_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_march_u8 -f python bench.py --workload single_1080p_u8 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_march_u8.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/*.ncu-rep | tail -8
