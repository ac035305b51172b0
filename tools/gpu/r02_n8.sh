#!/bin/bash
# Round 2, all 8 GPUs of one box (charged 8 x): the default bench line under torchrun, as the driver launches it.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu/r02_n8.sh 8'
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
( time timeout 420 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_default_n$N.log 2>&1; echo rc=$?; grep '^{' gpurun_out/bench_default_n$N.log | cut -c1-300
timeout 120 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --repeat 10 --timeline gpurun_out/timeline_8k_n$N.csv > gpurun_out/rowband_peer_8k_n$N.log 2>&1; tail -1 gpurun_out/rowband_peer_8k_n$N.log
