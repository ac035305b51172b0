N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
timeout 120 $TR tests/run_rowband_nccl.py --driver peer 2>&1 | grep -v Warning | tail -3
timeout 180 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -2
timeout 180 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 --mode exact 2>&1 | tail -2
timeout 180 $TR tests/run_rowband_nccl.py --driver nccl --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -2
