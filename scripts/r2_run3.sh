set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
for ilp in 1 2 4; do
CITAM_ACC_ILP=$ilp $B > gpurun_out/r3_ilp$ilp.json 2> gpurun_out/r3_ilp$ilp.err; echo rc=$?; tail -2 gpurun_out/r3_ilp$ilp.err | cut -c1-400
done
for ch in 96 192 256; do
$B --chunk $ch > gpurun_out/r3_c$ch.json 2> gpurun_out/r3_c$ch.err; echo rc=$?; tail -2 gpurun_out/r3_c$ch.err | cut -c1-400
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.3e"%d["value"], round(d["ms_per_step"],1), round(d["day"]["run_ms_per_day"],1), d["roofline"]["kernels_ms"])
    except Exception as e: print(f, e)
PY
