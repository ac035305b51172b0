#!/bin/bash
# Multi-GPU runs on one box:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu/scale.sh N'
#   batch sharding (default workload, pyramidal 4K), row bands over NVLink peer memory (8K), and the
#   bit-equality check of the row-band result against the single-GPU driver (fast and exact mode)
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_default_n$N.log; cut -c1-260 gpurun_out/bench_default_n$N.log
timeout 300 $TR bench.py --gpus $N --workload pyramidal_4k --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_pyr4k_n$N.log; cut -c1-260 gpurun_out/bench_pyr4k_n$N.log
timeout 300 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_n$N.log; cut -c1-260 gpurun_out/bench_8k_peer_n$N.log
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -1 > gpurun_out/rowband_peer_8k_n$N.log; cat gpurun_out/rowband_peer_8k_n$N.log
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --mode exact 2>&1 | tail -1 > gpurun_out/rowband_peer_8k_exact_n$N.log; cat gpurun_out/rowband_peer_8k_exact_n$N.log
