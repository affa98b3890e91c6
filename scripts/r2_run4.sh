set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
for cfg in "2 16" "4 4" "4 8" "4 16" "4 64" "2 64"; do
set -- $cfg
CITAM_ACC_ILP=$1 CITAM_ACC_BLOCKS=$2 $B > gpurun_out/r4_ilp$1_b$2.json 2> gpurun_out/r4_ilp$1_b$2.err; echo rc=$?
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r4_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.3e"%d["value"], round(d["ms_per_step"],1), round(d["day"]["run_ms_per_day"],1), d["roofline"]["kernels_ms"])
    except Exception as e: print(f, e)
PY
