set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize > gpurun_out/r7_a.json 2> gpurun_out/r7_a.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r7_a.json'));print('%.3e'%d['value'], d['ms_per_step'], d['roofline']['kernels_ms'], d['roofline']['frac'], d['roofline']['kernel'])"
python scripts/r2_finalize_probe.py 30
python scripts/r2_finalize_probe.py 32
