set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r10_bench.json 2> gpurun_out/r10_bench.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r10_bench.json'));print('%.3e'%d['value'], '%.3e'%d['e2e']['value'], round(d['ms_per_step'],2), d['roofline'], d['finalize'], d['e2e_with_results'])"
