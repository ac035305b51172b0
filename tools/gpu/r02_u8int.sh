#!/bin/bash
# Round 2: uint8 flavour with the frame sum / difference in the integer domain (default library) against the float-domain form (variant)
set -x
for i in 1 2; do
for lib in libof_b200.so libof_b200_u8f.so; do
  OF_B200_LIB_NAME=$lib timeout 200 python bench.py --workload single_1080p_u8 --workloads none --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_u8_$lib.$i.log 2>&1; echo "$lib rc=$?"; grep '^{' gpurun_out/bench_u8_$lib.$i.log | cut -c1-200
done
done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "uint8" > gpurun_out/pytest_u8.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_u8.log
