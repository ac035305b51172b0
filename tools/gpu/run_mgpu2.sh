N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
timeout 120 $TR tests/run_rowband_nccl.py --driver peer 2>&1 | tail -1
timeout 240 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_graph_n$N.log; cut -c1-330 gpurun_out/bench_8k_peer_graph_n$N.log
OF_B200_GRAPH=0 timeout 240 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_nograph_n$N.log; cut -c1-330 gpurun_out/bench_8k_peer_nograph_n$N.log
timeout 240 python bench.py --gpus 1 --workload pyramidal_8k --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_8k_n1.log; cut -c1-330 gpurun_out/bench_8k_n1.log
