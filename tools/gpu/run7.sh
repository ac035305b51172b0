set -x
timeout 300 python -m pytest tests -m gpu -x -q -k "fixed or timeout or smoke" 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload fixed_1080p --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fixed_march.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/bench_fixed_march.log") if x.startswith("{")]
print(open("gpurun_out/bench_fixed_march.log").read()[-800:] if not l else (lambda d:(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["parity"]))(json.loads(l[-1])))
PY
OF_B200_FIXED=tile python bench.py --workload fixed_1080p --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
