N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
timeout 300 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_repl_n$N.log; cut -c1-200 gpurun_out/bench_8k_peer_repl_n$N.log
OF_B200_ROWBAND_REPL_PX=0 timeout 300 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_norepl_n$N.log; cut -c1-200 gpurun_out/bench_8k_peer_norepl_n$N.log
OF_B200_ROWBAND_REPL_PX=2200000 timeout 300 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_repl2m_n$N.log; cut -c1-200 gpurun_out/bench_8k_peer_repl2m_n$N.log
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -1 > gpurun_out/rowband_peer_8k_repl_n$N.log; cat gpurun_out/rowband_peer_8k_repl_n$N.log
timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_default_n$N.log; cut -c1-200 gpurun_out/bench_default_n$N.log
