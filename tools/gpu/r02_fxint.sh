#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fixed or uint8 or rtl" > gpurun_out/pytest_fx.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_fx.log
for wl in fixed_1080p single_1080p_u8; do
  timeout 200 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$wl.log 2>&1; echo "$wl rc=$?"; grep '^{' gpurun_out/bench_$wl.log | cut -c1-200
done
