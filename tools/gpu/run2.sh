set -x
python bench.py --workload pyramidal_4k --batch 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b4_plan.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b4_plan.log | cut -c1-200
python bench.py --workload pyramidal_4k --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pyr_b16_plan.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_pyr_b16_plan.log | cut -c1-200
for nb in 1 2 3 4; do
OF_B200_BANDS=$nb python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_single_bands$nb.log 2>&1; echo rc=$?; tail -1 gpurun_out/bench_single_bands$nb.log | cut -c1-160
done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_single_auto.log 2>&1; tail -1 gpurun_out/bench_single_auto.log | cut -c1-160
ncu --set full --clock-control none --import-source on -k regex:pyramid_march --launch-skip 4 --launch-count 1 -o gpurun_out/prof_pyrmarch -f python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pyrmarch.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:warp_rows --launch-skip 24 --launch-count 1 -o gpurun_out/prof_warprows -f python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_warprows.log 2>&1; echo ncu rc=$?
ncu --set full --clock-control none --import-source on -k regex:upsample --launch-skip 5 --launch-count 1 -o gpurun_out/prof_upsample -f python bench.py --workload pyramidal_4k --batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_upsample.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/*.ncu-rep
