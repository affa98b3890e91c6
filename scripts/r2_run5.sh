set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
for ch in 48 64 96 128; do
$B --chunk $ch > gpurun_out/r5_c$ch.json 2> gpurun_out/r5_c$ch.err; echo rc=$?
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r5_*.json")):
    try:
        d=json.load(open(f))
        print(f, "%.3e"%d["value"], round(d["ms_per_step"],1), round(d["day"]["run_ms_per_day"],1), d["roofline"]["kernels_ms"])
    except Exception as e: print(f, e)
PY
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize"
$CMD > gpurun_out/r5_prof_plain.json 2> gpurun_out/r5_prof_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_expand_dyn|k_scan|k_scatter|k_select|k_pairs|k_accumulate' -s 1000 -c 10 -o gpurun_out/r5_prof_chunk -f $CMD > gpurun_out/r5_ncu.log 2>&1
echo full rc=$?
