set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wlname in single_1080p_u8 single_4k; do
python bench.py --workload $wlname --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wlname.log 2>&1; echo rc=$?; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_$wlname.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("$wlname", d["ms_per_step"], d["value"], d["roofline"]["frac"], d["parity"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
else: print(open("gpurun_out/bench_$wlname.log").read()[-1500:])
PY
done
