#!/bin/bash
# Exact-mode variants (split refinement + lk_tile5_kernel) through the whole GPU suite and the two exact benches:
#   gpurun --timeout 100 -- 'bash tools/gpu/exact_v2.sh'
export OF_B200_EXACT_REFINE=${OF_B200_EXACT_REFINE:-split} OF_B200_TILE=${OF_B200_TILE:-v2}
timeout 55 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_exact_v2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_exact_v2.log | cut -c1-400
for wl in single_1080p_exact pyramidal_4k_exact; do
  timeout 20 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_v2.log 2>&1; echo "$wl rc=$?"
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/bench_${wl}_v2.log") if x.startswith("{")][-1]; d=json.loads(l); print(d["config"]["name"], d["ms_per_step"], d["value"], d["parity"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/bench_${wl}_v2.log").read()[-600:])
PY
done
