N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
set -x
timeout 240 $TR bench.py --gpus $N --workload pyramidal_8k --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8k_peer_lanes2_n$N.log; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_8k_peer_lanes2_n$N.log") if x.startswith("{")]
print(open("gpurun_out/bench_8k_peer_lanes2_n$N.log").read()[-1500:] if not l else (lambda d:("lanes2", d["ms_per_step"], d["value"], d["parity"].get("rowband_bit_equal_to_single_gpu")))(json.loads(l[-1])))
PY
timeout 200 $TR tests/run_rowband_nccl.py --driver peer --height 4320 --width 7680 --levels 5 --iters 10 2>&1 | tail -1 > gpurun_out/rowband_peer_8k_final_n$N.log; cat gpurun_out/rowband_peer_8k_final_n$N.log
