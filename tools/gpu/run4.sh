set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for mb in 16 32 64 128; do
OF_B200_CHUNK_MB=$mb python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single_chunk$mb.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_single_chunk$mb.log") if x.startswith("{")][-1]
d=json.loads(l); print("chunk $mb", d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
done
