# the whole -m gpu suite + smoke + a short default bench (run under gpurun)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo rc=$?
tail -4 gpurun_out/bench_quick.err | cut -c1-500
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("%.3e e2e %.3e"%(d["value"], d["e2e"]["value"]), round(d["ms_per_step"],1), d["day"], d["roofline"]["kernels_ms"], d["finalize"], d["e2e_with_results"])
PY
