set -x
python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-finalize > gpurun_out/r16.json 2> gpurun_out/r16.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r16.json'));print('%.3e'%d['value'], round(d['ms_per_step'],2), d['roofline']['kernels_ms'], d.get('binned_fraction_of_agent_steps'))"
