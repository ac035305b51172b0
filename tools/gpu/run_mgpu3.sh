N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
timeout 240 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_tail_n$N.log; cut -c1-200 gpurun_out/bench_8k_peer_tail_n$N.log
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -1 > gpurun_out/rowband_peer_8k_tail_n$N.log; cat gpurun_out/rowband_peer_8k_tail_n$N.log
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --mode exact 2>&1 | tail -1
