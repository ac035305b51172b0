#!/bin/bash
# Round 2, N GPUs of one box:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu/r02_n2.sh N'
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_default_n$N.log 2>&1; echo rc=$?; grep '^{' gpurun_out/bench_default_n$N.log | cut -c1-300
# NVLink byte counters of GPU 0 around one row-band job (8K, 5 x 10, fast): the peer stores of the kernels are the only NVLink traffic
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before_n$N.txt 2>&1
timeout 300 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --repeat 20 --timeline gpurun_out/timeline_8k_n$N.csv > gpurun_out/rowband_peer_8k_n$N.log 2>&1; tail -1 gpurun_out/rowband_peer_8k_n$N.log
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_n$N.txt 2>&1
timeout 300 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --mode exact > gpurun_out/rowband_peer_8k_exact_n$N.log 2>&1; tail -1 gpurun_out/rowband_peer_8k_exact_n$N.log
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_process or times_out or rowband" > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_n$N.log
