#!/bin/bash
# Round 2: ring depth / CTAs per SM of the float marching kernels (variant libraries) -- one GPU
set -x
for lib in libof_b200.so libof_b200_s2.so libof_b200_s2c3.so; do
  for wl in single_1080p pyramidal_4k; do
    OF_B200_LIB_NAME=$lib timeout 300 python bench.py --workload $wl --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_$lib.log 2>&1; echo "$lib $wl rc=$?"; grep '^{' gpurun_out/bench_${wl}_$lib.log | cut -c1-200
  done
done
